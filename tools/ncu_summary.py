"""Summarise an .ncu-rep (read here, no GPU): headline metrics + stall reasons + hottest SASS lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__cluster_size"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print("  %-70s %-14s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si, ii = h.index("# Samples"), h.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) == len(h) and r[si].isdigit()]      # several kernels: their source pages follow one another
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "sass instructions", len(data))
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[h.index(c)]) for r in data) for c in stalls}
for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-26s %9d %.3f" % (c, v, v / max(tot, 1)))
top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:topn]
for i in sorted(top):
    r = data[i]
    best = max(stalls, key=lambda c: int(r[h.index(c)]))
    print("%6d samples=%7s exec=%10s %-16s %s" % (i, r[si], r[ii], best, r[1][:80]))
