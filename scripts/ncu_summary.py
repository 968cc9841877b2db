"""Summarise an .ncu-rep (raw page + hottest SASS lines) into text for profiles/."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform"]
print(f"# {rep}")
for h, u, v in zip(hdr, units, vals):
    if any(h == w or h.startswith(w + ".") and len(h) - len(w) < 24 for w in want) or h in want:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
i_s, i_n, i_e = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[i_n]), int(r[i_e]), r[i_s], r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print(f"\n# hottest SASS (of {tot} warp samples, {len(data)} instructions)")
for d in sorted(data, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    st = sorted(((hdr[i], int(d[3][i] or 0)) for i in stall), key=lambda x: -x[1])[:2]
    print(f"{100 * d[0] / tot:5.1f}%  exec={d[1]:>10d}  {d[2][:72]:72s} {st}")
