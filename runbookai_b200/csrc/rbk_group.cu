// rbk_group.cu — the multi-GPU index behind ONE handle and ONE call (include/rbk_knn.h, rbk_group_*).
//
// RunbookAI is a single Node process: the deployment that shards a corpus over the GPUs of a
// box is not "one process per GPU under torchrun" (that is bench.py's harness) but one host thread calling
// rbk_group_search_f32.  A group owns one rbk_index per GPU, deals rows out block-cyclically (global slot s lives
// on device (s / block) % G - the index grows at sync time, so no device needs the final corpus size), and answers
// a batch with
//     H2D of the queries to every GPU  ->  the enqueue-only fused scan + finalize on every GPU (its own stream)
//     ->  ONE ncclAllGather of the packed per-GPU blocks (results + exactness flags) over NVLink
//     ->  merge kernel on GPU 0  ->  one D2H, one host synchronisation.
// NCCL is resolved with dlopen("libnccl.so.2") on first use, so the library itself has no link-time dependency on
// it and a one-GPU group never needs it.  Exact per-shard fp64 scores make the merge exact; local row order is
// global slot order within a device, so the (score desc, slot asc) tie-break survives (SlotLayout, rbk_internal.h).
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: every NCCL symbol is looked up at run time
#include <string.h>

#include <algorithm>
#include <memory>

#include "rbk_index_impl.h"

using namespace rbk;
using namespace rbk::impl;

namespace {

struct NcclApi {
  void* handle = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  std::string error;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?");
      return;
    }
    auto sym = [&](const char* n) { return dlsym(api.handle, n); };
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.ok = api.GetErrorString && api.CommInitAll && api.CommDestroy && api.AllGather && api.GroupStart && api.GroupEnd;
    if (!api.ok) api.error = "libnccl.so.2 lacks a required symbol";
  });
  return api;
}

rbk_status nccl_fail(ncclResult_t r, const char* what) {
  NcclApi& n = nccl_api();
  return fail(RBK_ENCCL, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(r) : "NCCL error"));
}
#define NC(expr)                                       \
  do {                                                 \
    ncclResult_t _r = (expr);                          \
    if (_r != ncclSuccess) return nccl_fail(_r, #expr); \
  } while (0)

}  // namespace

struct rbk_group {
  int dim = 0, G = 0;
  int64_t block = 4096;          // rows per placement block
  int64_t n_slots = 0;           // global slots handed out (tombstones included)
  std::vector<int> devices;
  std::vector<rbk_index*> parts;
  std::vector<ncclComm_t> comms; // empty for G == 1
  std::mutex mu;
  struct Dev {
    DevBuf<unsigned char> q, local, all;
  };
  std::vector<Dev> dev;
  DevBuf<unsigned char> out;     // device 0: slots | scores | counts | flags[B+1]
  PinBuf<unsigned char> h_out, h_q;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t redone_batches = 0;
};

namespace {

// global slot -> (device, local row)
inline void locate(const rbk_group* g, int64_t slot, int* dev, int64_t* local) {
  const int64_t blk = slot / g->block;
  *dev = static_cast<int>(blk % g->G);
  *local = (blk / g->G) * g->block + slot % g->block;
}

rbk_status group_append(rbk_group* g, const void* rows, int elem, int64_t n, int64_t* first_out) {
  if (!g) return fail(RBK_EINVAL, "null group");
  if (n < 0 || (n > 0 && !rows)) return fail(RBK_EINVAL, "bad rows argument");
  std::lock_guard<std::mutex> lk(g->mu);
  if (first_out) *first_out = g->n_slots;
  const size_t row_bytes = static_cast<size_t>(g->dim) * elem;
  int64_t done = 0;
  while (done < n) {
    const int64_t s = g->n_slots;
    const int64_t take = std::min<int64_t>(n - done, g->block - s % g->block);   // up to the end of this block
    int d;
    int64_t local;
    locate(g, s, &d, &local);
    const unsigned char* src = static_cast<const unsigned char*>(rows) + static_cast<size_t>(done) * row_bytes;
    int64_t first = -1;
    rbk_status st = elem == 8   ? rbk_index_append_f64(g->parts[d], reinterpret_cast<const double*>(src), take, &first)
                    : elem == 4 ? rbk_index_append_f32(g->parts[d], reinterpret_cast<const float*>(src), take, &first)
                                : rbk_index_append_bf16(g->parts[d], reinterpret_cast<const uint16_t*>(src), take, &first);
    if (st != RBK_OK) return st;
    if (first != local) return fail(RBK_EINVAL, "group placement out of step with a member index");
    g->n_slots += take;
    done += take;
  }
  return RBK_OK;
}

// Split global slots (and optionally their rows) by owning device.
void split_slots(const rbk_group* g, const int64_t* slots, int64_t n, std::vector<std::vector<int64_t>>* local,
                 std::vector<std::vector<int64_t>>* order) {
  local->assign(g->G, {});
  if (order) order->assign(g->G, {});
  for (int64_t i = 0; i < n; ++i) {
    int d;
    int64_t l;
    locate(g, slots[i], &d, &l);
    (*local)[d].push_back(l);
    if (order) (*order)[d].push_back(i);
  }
}

// all-gather of the packed blocks (G > 1) + merge on device 0 into g->out; enqueue only
rbk_status exchange_and_merge(rbk_group* g, int B, int k_fetch, size_t blk) {
  if (g->G > 1) {
    NcclApi& n = nccl_api();
    NC(n.GroupStart());
    for (int d = 0; d < g->G; ++d) {
      ncclResult_t r = n.AllGather(g->dev[d].local.p, g->dev[d].all.p, blk, ncclUint8, g->comms[d], g->parts[d]->stream);
      if (r != ncclSuccess) {
        n.GroupEnd();
        return nccl_fail(r, "ncclAllGather");
      }
    }
    NC(n.GroupEnd());
  }
  DeviceGuard dg(g->devices[0]);
  const size_t nk = static_cast<size_t>(B) * k_fetch;
  const char* base = reinterpret_cast<const char*>(g->G > 1 ? g->dev[0].all.p : g->dev[0].local.p);
  unsigned char* o = g->out.p;
  CK(launch_merge_shards(g->G, B, k_fetch, base, base + nk * 8, base + nk * 16,
                         base + rbk_packed_flags_offset(B, k_fetch), blk, blk, blk, blk,
                         reinterpret_cast<long long*>(o), reinterpret_cast<double*>(o + nk * 8),
                         reinterpret_cast<int*>(o + nk * 16), reinterpret_cast<int*>(o + nk * 16 + static_cast<size_t>(B) * 4),
                         g->parts[0]->stream));
  return RBK_OK;
}

rbk_status group_search(rbk_group* g, const void* queries, int elem, int32_t B, int32_t query_dim, int32_t k_fetch,
                        double min_score, int64_t* out_slots, double* out_scores, int32_t* out_counts, float* ms_out) {
  if (!g) return fail(RBK_EINVAL, "null group");
  if (B > 0 && (!out_slots || !out_scores || !out_counts)) return fail(RBK_EINVAL, "null output");
  rbk_status st = check_search_args(g->parts[0], B, queries != nullptr, query_dim, k_fetch, min_score);
  if (st != RBK_OK) return st;
  if (ms_out) *ms_out = 0.f;
  if (B == 0) return RBK_OK;
  std::lock_guard<std::mutex> lk(g->mu);
  const size_t q_bytes = static_cast<size_t>(B) * g->dim * elem;
  const size_t blk = static_cast<size_t>(rbk_packed_block_bytes(B, k_fetch));
  const size_t off_f = static_cast<size_t>(rbk_packed_flags_offset(B, k_fetch));
  const size_t nk = static_cast<size_t>(B) * k_fetch;
  const size_t out_bytes = nk * 16 + (2 * static_cast<size_t>(B) + 1) * 4;
  {
    DeviceGuard dg(g->devices[0]);
    CK(g->h_q.ensure(q_bytes));
    CK(g->h_out.ensure(out_bytes));
    CK(g->out.ensure(out_bytes));
    // the merge kernel ADDS the number of dirty queries to the word after the flags: start every search from zero
    // (its position depends on B and k_fetch, so a running count across calls of different shapes would be garbage)
    CK(cudaMemsetAsync(g->out.p + out_bytes - 4, 0, 4, g->parts[0]->stream));
  }
  memcpy(g->h_q.p, queries, q_bytes);   // pinned staging: the G H2D copies below run concurrently, one per PCIe link
  const int src_type = elem == 8 ? 0 : 1;
  for (int d = 0; d < g->G; ++d) {
    rbk_index* ix = g->parts[d];
    std::lock_guard<std::mutex> il(ix->mu);
    DeviceGuard dg(ix->device);
    CK(g->dev[d].q.ensure(q_bytes));
    CK(g->dev[d].local.ensure(blk));
    if (g->G > 1) CK(g->dev[d].all.ensure(blk * g->G));
    if (d == 0) CK(cudaEventRecord(g->ev0, ix->stream));
    CK(cudaMemcpyAsync(g->dev[d].q.p, g->h_q.p, q_bytes, cudaMemcpyHostToDevice, ix->stream));
    st = ensure_query_scratch(ix, B, elem);
    if (st != RBK_OK) return st;
    unsigned char* l = g->dev[d].local.p;
    st = enqueue_search(ix, g->dev[d].q.p, src_type, B, k_fetch, min_score, reinterpret_cast<long long*>(l),
                        reinterpret_cast<double*>(l + nk * 8), reinterpret_cast<int*>(l + nk * 16),
                        reinterpret_cast<int*>(l + off_f));
    if (st != RBK_OK) return st;
  }
  st = exchange_and_merge(g, B, k_fetch, blk);
  if (st != RBK_OK) return st;
  rbk_index* i0 = g->parts[0];
  {
    DeviceGuard dg(i0->device);
    CK(cudaMemcpyAsync(g->h_out.p, g->out.p, out_bytes, cudaMemcpyDeviceToHost, i0->stream));
    CK(cudaEventRecord(g->ev1, i0->stream));
    CK(cudaStreamSynchronize(i0->stream));   // the ONE host round trip of an exact batch
  }
  int dirty_total = 0;
  memcpy(&dirty_total, g->h_out.p + out_bytes - 4, 4);
  if (dirty_total != 0) {
    // some shard could not prove a query (more near-ties than its candidate margin): every shard re-answers the
    // batch through the synchronous path (wide rescan, then the exhaustive fp64 kernel), and the exchange is redone
    g->redone_batches++;
    for (int d = 0; d < g->G; ++d) {
      unsigned char* l = g->dev[d].local.p;
      st = rbk_index_search_device(g->parts[d], g->dev[d].q.p, B, k_fetch, min_score, l, l + nk * 8, l + nk * 16);
      if (st != RBK_OK) return st;
      DeviceGuard dg(g->parts[d]->device);
      CK(cudaMemsetAsync(l + off_f, 0, static_cast<size_t>(B) * 4, g->parts[d]->stream));   // exact by construction
    }
    st = exchange_and_merge(g, B, k_fetch, blk);
    if (st != RBK_OK) return st;
    DeviceGuard dg(i0->device);
    CK(cudaMemcpyAsync(g->h_out.p, g->out.p, out_bytes, cudaMemcpyDeviceToHost, i0->stream));
    CK(cudaEventRecord(g->ev1, i0->stream));
    CK(cudaStreamSynchronize(i0->stream));
  }
  if (ms_out) cudaEventElapsedTime(ms_out, g->ev0, g->ev1);
  memcpy(out_slots, g->h_out.p, nk * 8);
  memcpy(out_scores, g->h_out.p + nk * 8, nk * 8);
  memcpy(out_counts, g->h_out.p + nk * 16, static_cast<size_t>(B) * 4);
  return RBK_OK;
}

}  // namespace

extern "C" {

rbk_status rbk_group_create(int32_t dim, const int32_t* device_ids, int32_t n_devices, int64_t capacity_hint,
                            uint32_t flags, rbk_group** out) {
  if (!out) return fail(RBK_EINVAL, "out is null");
  *out = nullptr;
  if (!device_ids || n_devices < 1 || n_devices > 64) return fail(RBK_EINVAL, "bad device list");
  for (int a = 0; a < n_devices; ++a)
    for (int b = a + 1; b < n_devices; ++b)
      if (device_ids[a] == device_ids[b]) return fail(RBK_EINVAL, "a device may appear only once in a group");
  std::unique_ptr<rbk_group> g(new (std::nothrow) rbk_group());
  if (!g) return fail(RBK_ENOMEM, "out of host memory");
  g->dim = dim;
  g->G = n_devices;
  g->devices.assign(device_ids, device_ids + n_devices);
  g->dev.resize(n_devices);
  auto destroy_parts = [&]() {
    for (rbk_index* ix : g->parts) rbk_index_destroy(ix);
    g->parts.clear();
  };
  for (int d = 0; d < n_devices; ++d) {
    rbk_index* ix = nullptr;
    rbk_status st = rbk_index_create_ex(dim, device_ids[d], capacity_hint / n_devices + g->block, flags, &ix);
    if (st != RBK_OK) {
      destroy_parts();
      return st;
    }
    ix->slot.block = static_cast<int32_t>(g->block);
    ix->slot.G = n_devices;
    ix->slot.g = d;
    g->parts.push_back(ix);
  }
  if (n_devices > 1) {
    NcclApi& n = nccl_api();
    if (!n.ok) {
      destroy_parts();
      return fail(RBK_ENCCL, n.error + " (a group of more than one GPU exchanges its per-GPU lists with ncclAllGather)");
    }
    g->comms.resize(n_devices);
    ncclResult_t r = n.CommInitAll(g->comms.data(), n_devices, g->devices.data());
    if (r != ncclSuccess) {
      g->comms.clear();
      destroy_parts();
      return nccl_fail(r, "ncclCommInitAll");
    }
  }
  {
    DeviceGuard dg(g->devices[0]);
    if (cudaEventCreate(&g->ev0) != cudaSuccess || cudaEventCreate(&g->ev1) != cudaSuccess) {
      rbk_group_destroy(g.release());
      return fail(RBK_ECUDA, "cudaEventCreate");
    }
  }
  *out = g.release();
  return RBK_OK;
}

void rbk_group_destroy(rbk_group* g) {
  if (!g) return;
  for (size_t d = 0; d < g->parts.size(); ++d) {
    DeviceGuard dg(g->devices[d]);
    if (g->parts[d] && g->parts[d]->stream) cudaStreamSynchronize(g->parts[d]->stream);
    g->dev[d].q.release();
    g->dev[d].local.release();
    g->dev[d].all.release();
  }
  if (!g->comms.empty()) {
    NcclApi& n = nccl_api();
    for (ncclComm_t c : g->comms)
      if (c && n.CommDestroy) n.CommDestroy(c);
  }
  if (!g->devices.empty()) {
    DeviceGuard dg(g->devices[0]);
    g->out.release();
    g->h_out.release();
    g->h_q.release();
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
  }
  for (rbk_index* ix : g->parts) rbk_index_destroy(ix);
  delete g;
}

rbk_status rbk_group_append_f64(rbk_group* g, const double* rows, int64_t n, int64_t* first) {
  return group_append(g, rows, 8, n, first);
}
rbk_status rbk_group_append_f32(rbk_group* g, const float* rows, int64_t n, int64_t* first) {
  return group_append(g, rows, 4, n, first);
}
rbk_status rbk_group_append_bf16(rbk_group* g, const uint16_t* rows, int64_t n, int64_t* first) {
  return group_append(g, rows, 2, n, first);
}

rbk_status rbk_group_overwrite_f64_batch(rbk_group* g, const int64_t* slots, int64_t n, const double* rows) {
  if (!g) return fail(RBK_EINVAL, "null group");
  if (n < 0 || (n > 0 && (!slots || !rows))) return fail(RBK_EINVAL, "bad argument");
  std::lock_guard<std::mutex> lk(g->mu);
  for (int64_t i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= g->n_slots) return fail(RBK_EINVAL, "slot out of range");
  std::vector<std::vector<int64_t>> local, order;
  split_slots(g, slots, n, &local, &order);
  rbk_status worst = RBK_OK;
  std::string msg;
  for (int d = 0; d < g->G; ++d) {
    if (local[d].empty()) continue;
    std::vector<double> part(local[d].size() * static_cast<size_t>(g->dim));
    for (size_t i = 0; i < order[d].size(); ++i)
      memcpy(&part[i * g->dim], rows + static_cast<size_t>(order[d][i]) * g->dim, sizeof(double) * g->dim);
    rbk_status st = rbk_index_overwrite_f64_batch(g->parts[d], local[d].data(), static_cast<int64_t>(local[d].size()),
                                                  part.data());
    if (st != RBK_OK) {   // keep going: the live slots of the other devices are still written, as within one index
      worst = st;
      msg = last_error();
    }
  }
  return worst == RBK_OK ? RBK_OK : fail(worst, msg);
}

rbk_status rbk_group_tombstone(rbk_group* g, const int64_t* slots, int64_t n) {
  if (!g) return fail(RBK_EINVAL, "null group");
  if (n < 0 || (n > 0 && !slots)) return fail(RBK_EINVAL, "bad slots argument");
  std::lock_guard<std::mutex> lk(g->mu);
  for (int64_t i = 0; i < n; ++i)
    if (slots[i] < 0 || slots[i] >= g->n_slots) return fail(RBK_EINVAL, "slot out of range");
  std::vector<std::vector<int64_t>> local;
  split_slots(g, slots, n, &local, nullptr);
  for (int d = 0; d < g->G; ++d) {
    if (local[d].empty()) continue;
    rbk_status st = rbk_index_tombstone(g->parts[d], local[d].data(), static_cast<int64_t>(local[d].size()));
    if (st != RBK_OK) return st;
  }
  return RBK_OK;
}

rbk_status rbk_group_clear(rbk_group* g) {
  if (!g) return fail(RBK_EINVAL, "null group");
  std::lock_guard<std::mutex> lk(g->mu);
  for (rbk_index* ix : g->parts) {
    rbk_status st = rbk_index_clear(ix);
    if (st != RBK_OK) return st;
  }
  g->n_slots = 0;
  return RBK_OK;
}

int64_t rbk_group_count(const rbk_group* g) {
  int64_t n = 0;
  if (g)
    for (const rbk_index* ix : g->parts) n += rbk_index_count(ix);
  return n;
}
int64_t rbk_group_size(const rbk_group* g) { return g ? g->n_slots : 0; }
int32_t rbk_group_devices(const rbk_group* g) { return g ? g->G : 0; }
rbk_index* rbk_group_member(rbk_group* g, int32_t i) { return (g && i >= 0 && i < g->G) ? g->parts[i] : nullptr; }
int64_t rbk_group_redone_batches(const rbk_group* g) { return g ? g->redone_batches : 0; }

rbk_status rbk_group_search_f32(rbk_group* g, const float* queries, int32_t B, int32_t query_dim, int32_t k_fetch,
                                double min_score, int64_t* out_slots, double* out_scores, int32_t* out_counts,
                                float* device_ms_out) {
  return group_search(g, queries, 4, B, query_dim, k_fetch, min_score, out_slots, out_scores, out_counts, device_ms_out);
}
rbk_status rbk_group_search_f64(rbk_group* g, const double* queries, int32_t B, int32_t query_dim, int32_t k_fetch,
                                double min_score, int64_t* out_slots, double* out_scores, int32_t* out_counts,
                                float* device_ms_out) {
  return group_search(g, queries, 8, B, query_dim, k_fetch, min_score, out_slots, out_scores, out_counts, device_ms_out);
}

}  // extern "C"
