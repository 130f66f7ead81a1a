"""Turn the raw ncu outputs in gpurun_out/ into the small committed summaries under profiles/.
usage: python profiles/summarise.py <tag> <launch csv> <ncu-rep>"""
import csv
import gzip
import shutil
import subprocess
import sys

tag, launches, rep = sys.argv[1:4]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
total = sum(sum(v) for v in agg.values())
with open("profiles/%s_launch_summary.txt" % tag, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches:\n"
            "# compare SHARES, not absolutes).  kernel | launches | total ms | share | mean us\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-90s %5d %10.3f %6.1f%% %10.1f\n" % (k[:90], len(v), sum(v) / 1e6, 100 * sum(v) / total, sum(v) / len(v) / 1e3))
    f.write("total %.3f ms over %d launches\n" % (total / 1e6, sum(len(v) for v in agg.values())))
with open(launches, "rb") as fi, gzip.open("profiles/%s_launches.csv.gz" % tag, "wb") as fo:
    shutil.copyfileobj(fi, fo)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
with open("profiles/%s_ncu_full_summary.txt" % tag, "w") as f:
    f.write("# ncu --set full --clock-control none, one launch per kernel (iteration 3 of a cfg-4 solve, 1 B200)\n")
    tot = 0.0
    for r in rr[2:]:
        f.write("== %s\n" % r[h.index("Kernel Name")])
        for k in keys:
            if k in h:
                f.write("   %-85s %s %s\n" % (k, r[h.index(k)], units[h.index(k)]))
        def val(k):
            v, u = float(r[h.index(k)]), units[h.index(k)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        t = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        tot += t
        f.write("   DRAM traffic (read+write) %.0f bytes\n" % t)
    f.write("DRAM traffic of one PGD iteration (both kernels): %.0f bytes\n" % tot)
print(open("profiles/%s_launch_summary.txt" % tag).read())
print(open("profiles/%s_ncu_full_summary.txt" % tag).read())
