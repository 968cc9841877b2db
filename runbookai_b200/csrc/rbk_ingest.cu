// rbk_ingest.cu — K3: corpus ingest.  Replaces VectorStore.loadEmbeddings /
// bufferToFloatArray (reference src/knowledge/store/vector-store.ts:56-88): rows arrive as
// little-endian float64 (the SQLite BLOB layout), float32 or bf16 and are stored as bf16
// rows of pitch dpad (= dim rounded up to 8 elements, zero padded) plus, per row,
//   inv_norm (fp32)  1/||row||           for the approximate scan (NaN = never matches)
//   norm2    (fp64)  sum of squares accumulated in index order, multiply-then-add — the
//                    reference's `normB` (embedder.ts:175-181) bit for bit, reused by the
//                    exact re-rank so it is not recomputed per query.
// Both kernels are HBM-bound streams: 16-byte vector loads/stores, no reuse.
#include <cuda_bf16.h>

#include "rbk_internal.h"

namespace rbk {

namespace {

__device__ __forceinline__ uint16_t f64_to_bf16_bits(double x) {
  // f64 -> f32 (RNE) -> bf16 (RNE); documented in DESIGN.md §3.
  return __bfloat16_as_ushort(__float2bfloat16_rn(__double2float_rn(x)));
}
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float x) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}

// One thread produces 8 consecutive output elements (one 16-byte store).
template <typename SrcT>
__global__ void __launch_bounds__(256) convert_rows_kernel(const SrcT* __restrict__ src, int64_t n_rows, int d,
                                                           int dpad, uint16_t* __restrict__ dst,
                                                           double* __restrict__ dst_f64, bool aligned) {
  const int groups = dpad >> 3;
  const int64_t total = n_rows * groups;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = g / groups;
    const int c0 = static_cast<int>(g - row * groups) << 3;
    const SrcT* s = src + row * d + c0;
    uint16_t o[8];
    if (aligned) {
      if constexpr (sizeof(SrcT) == 8) {
        const double2* s2 = reinterpret_cast<const double2*>(s);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 v = __ldg(s2 + j);
          o[2 * j] = f64_to_bf16_bits(v.x);
          o[2 * j + 1] = f64_to_bf16_bits(v.y);
        }
      } else if constexpr (sizeof(SrcT) == 4) {
        const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float4 v = __ldg(s4 + j);
          o[4 * j] = f32_to_bf16_bits(v.x);
          o[4 * j + 1] = f32_to_bf16_bits(v.y);
          o[4 * j + 2] = f32_to_bf16_bits(v.z);
          o[4 * j + 3] = f32_to_bf16_bits(v.w);
        }
      } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(s));
        *reinterpret_cast<uint4*>(o) = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint16_t b = 0;
        if (c0 + j < d) {
          if constexpr (sizeof(SrcT) == 8) b = f64_to_bf16_bits(static_cast<double>(s[j]));
          else if constexpr (sizeof(SrcT) == 4) b = f32_to_bf16_bits(static_cast<float>(s[j]));
          else b = static_cast<uint16_t>(s[j]);
        }
        o[j] = b;
      }
    }
    if (dst_f64 != nullptr) {   // exact-source sidecar: the original values, widened to f64 (pitch d)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < d) {
          double x;
          if constexpr (sizeof(SrcT) == 8) x = static_cast<double>(s[j]);
          else if constexpr (sizeof(SrcT) == 4) x = static_cast<double>(static_cast<float>(s[j]));
          else x = static_cast<double>(__uint_as_float(static_cast<uint32_t>(s[j]) << 16));
          dst_f64[row * d + c0 + j] = x;
        }
      }
    }
    uint4 out;
    out.x = o[0] | (static_cast<uint32_t>(o[1]) << 16);
    out.y = o[2] | (static_cast<uint32_t>(o[3]) << 16);
    out.z = o[4] | (static_cast<uint32_t>(o[5]) << 16);
    out.w = o[6] | (static_cast<uint32_t>(o[7]) << 16);
    *reinterpret_cast<uint4*>(dst + row * dpad + c0) = out;
  }
}

__device__ __forceinline__ double bf16_bits_to_f64(uint32_t h) {
  return static_cast<double>(__uint_as_float(h << 16));
}

// One thread per row: the accumulation order is part of the parity contract.
// rows_f64 == nullptr: the bf16 row IS the corpus row, norm2 is its exact sequential sum of squares.
// rows_f64 != nullptr: the f64 sidecar is the corpus row (norm2 from it); the bf16 row only feeds the
//   approximate scan, and the angle between the two (an upper bound on how far the approximate cosine of
//   this row can be from its true cosine, on top of the other error terms) is folded into *eps_c_max.
__global__ void __launch_bounds__(128) row_norms_kernel(const uint16_t* __restrict__ rows,
                                                        const double* __restrict__ rows_f64, int64_t n_rows, int d,
                                                        int dpad, float* __restrict__ inv_norm,
                                                        double* __restrict__ norm2, int* __restrict__ eps_c_max) {
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= n_rows) return;
  const uint4* p = reinterpret_cast<const uint4*>(rows + row * dpad);
  double acc = 0.0;
  for (int g = 0; g < (dpad >> 3); ++g) {  // pad columns are zero: adding 0*0 is exact
    const uint4 v = __ldg(p + g);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double lo = bf16_bits_to_f64(w[j] & 0xFFFFu);
      const double hi = bf16_bits_to_f64(w[j] >> 16);
      acc = __dadd_rn(acc, __dmul_rn(lo, lo));
      acc = __dadd_rn(acc, __dmul_rn(hi, hi));
    }
  }
  const bool ok = acc > 0.0 && acc < INFINITY;
  inv_norm[row] = ok ? static_cast<float>(1.0 / sqrt(acc)) : __uint_as_float(0x7FC00000u);
  if (rows_f64 == nullptr) {
    norm2[row] = acc;
    return;
  }
  const double* x = rows_f64 + row * d;
  const uint16_t* xb = rows + row * dpad;
  double n2 = 0.0, diff2 = 0.0;
  for (int i = 0; i < d; ++i) {
    const double v = x[i];
    n2 = __dadd_rn(n2, __dmul_rn(v, v));   // the reference's normB for the f64 row
    const double e = v - bf16_bits_to_f64(xb[i]);
    diff2 += e * e;
  }
  norm2[row] = n2;
  float eps = 0.f;
  if (diff2 > 0.0) {
    const double ratio = (n2 > 0.0 && n2 < INFINITY) ? sqrt(diff2 / n2) * (1.0 + 1e-9) : 2.0;
    eps = static_cast<float>((ratio < 1.0 ? asin(ratio) : 3.2) * (1.0 + 1e-6)) ;
    eps = nextafterf(eps, INFINITY);
  }
  if (eps > 0.f) atomicMax(eps_c_max, __float_as_int(eps));   // non-negative floats order like ints
}

__global__ void tombstone_kernel(const int64_t* __restrict__ slots, int64_t n, int64_t n_rows,
                                 float* __restrict__ inv_norm, unsigned int* __restrict__ dead_bits,
                                 int* __restrict__ n_killed) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int64_t s = slots[i];
  if (s < 0 || s >= n_rows) return;
  const unsigned int bit = 1u << (s & 31);
  const unsigned int old = atomicOr(dead_bits + (s >> 5), bit);
  if (!(old & bit)) {
    inv_norm[s] = __uint_as_float(0x7FC00000u);
    atomicAdd(n_killed, 1);
  }
}

int grid_for(int64_t items, int threads, int max_blocks) {
  int64_t b = (items + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

}  // namespace

cudaError_t launch_convert_rows(const void* src, int src_type, int64_t n_rows, int d, int dpad, uint16_t* dst_rows,
                                double* dst_f64, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  const int64_t total = n_rows * (dpad >> 3);
  const int grid = grid_for(total, 256, 148 * 16);
  // vector loads need d % 8 == 0 (every 8-group starts 16-byte aligned) and an aligned base
  const bool aligned = (d & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  if (src_type == 0)
    convert_rows_kernel<double>
        <<<grid, 256, 0, stream>>>(static_cast<const double*>(src), n_rows, d, dpad, dst_rows, dst_f64, aligned);
  else if (src_type == 1)
    convert_rows_kernel<float>
        <<<grid, 256, 0, stream>>>(static_cast<const float*>(src), n_rows, d, dpad, dst_rows, dst_f64, aligned);
  else
    convert_rows_kernel<uint16_t>
        <<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(src), n_rows, d, dpad, dst_rows, dst_f64, aligned);
  return cudaGetLastError();
}

cudaError_t launch_row_norms(const uint16_t* rows, const double* rows_f64, int64_t n_rows, int d, int dpad,
                             float* inv_norm, double* norm2, int* eps_c_max, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  const int64_t blocks = (n_rows + 127) / 128;
  row_norms_kernel<<<static_cast<unsigned>(blocks), 128, 0, stream>>>(rows, rows_f64, n_rows, d, dpad, inv_norm, norm2,
                                                                      eps_c_max);
  return cudaGetLastError();
}

cudaError_t launch_tombstone(const int64_t* dev_slots, int64_t n, int64_t n_rows, float* inv_norm,
                             unsigned int* dead_bits, int* n_killed, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  tombstone_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(dev_slots, n, n_rows, inv_norm,
                                                                              dead_bits, n_killed);
  return cudaGetLastError();
}

}  // namespace rbk
