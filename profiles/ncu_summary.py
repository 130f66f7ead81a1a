"""Key counters of every launch in an `ncu --set full` report -> a small text summary (committed under profiles/).
usage: python profiles/ncu_summary.py <report.ncu-rep> <out.txt> ["title line"] [--traffic-json out.json --sources a.cu,b.cuh]

With --traffic-json it also writes {kernel: dram bytes per launch} together with the sha1 of the listed kernel
sources, so that bench.py can refuse a capture taken on other code (roofline.traffic stays null then)."""
import csv
import hashlib
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    args = sys.argv[1:]
    tj, srcs = None, []
    if "--traffic-json" in args:
        i = args.index("--traffic-json")
        tj = args[i + 1]
        del args[i:i + 2]
    if "--sources" in args:
        i = args.index("--sources")
        srcs = args[i + 1].split(",")
        del args[i:i + 2]
    rep, out = args[0], args[1]
    title = args[2] if len(args) > 2 else "ncu --set full --clock-control none"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    traffic = {}
    with open(out, "w") as f:
        f.write("# %s\n" % title)
        for r in rr[2:]:
            name = r[h.index("Kernel Name")]
            f.write("== %s\n" % name)
            for k in KEYS:
                if k in h:
                    f.write("   %-90s %s %s\n" % (k, r[h.index(k)], units[h.index(k)]))
            t = sum(float(r[h.index(k)]) * SCALE[units[h.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            f.write("   DRAM traffic (read+write) %.0f bytes\n" % t)
            traffic.setdefault(name.split("(")[0], []).append(t)
    if tj:
        sha = hashlib.sha1()
        for s in srcs:
            sha.update(open(s, "rb").read())
        json.dump({"report": os.path.basename(rep), "sources": srcs, "sources_sha1": sha.hexdigest(),
                   "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}}, open(tj, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main()
