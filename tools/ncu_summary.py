#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): key raw metrics + per-opcode and per-source-line
instruction shares.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv, io, subprocess, sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    h, units, vals = rows[0], rows[1], rows[2]
    print("kernel:", vals[h.index("Kernel Name")], file=out)
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print("%-90s %s %s" % (k, vals[i], units[i]), file=out)
    srows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hdr = srows[1]; data = srows[2:]
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iex]) for r in data)
    print("\nwarp instructions executed: %d  (static SASS lines %d)" % (tot, len(data)), file=out)
    c = Counter()
    for r in data:
        op = r[isrc].split()
        if not op:
            continue
        o = op[1] if op[0].startswith("@") else op[0]
        c[o.split(".")[0]] += int(r[iex])
    print("opcode shares:", ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in c.most_common(18)), file=out)
    crows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    print("\nsource lines by executed warp instructions (top 45; second column = share of stall samples):", file=out)
    best = []
    cur_file = None
    hdr2 = None
    for r in crows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1]; continue
        if len(r) > 3 and r[0] == "Line No":
            hdr2 = r; continue
        if hdr2 and len(r) == len(hdr2) and r[2] == "-":
            try:
                n = int(r[hdr2.index("Instructions Executed")]); ss = int(r[hdr2.index("# Samples")])
            except ValueError:
                continue
            if n:
                best.append((n, ss, cur_file.split("/")[-1], r[0], r[1].strip()[:105]))
    tots = sum(x[1] for x in best) or 1
    best.sort(reverse=True)
    for n, ss, f, ln, src in best[:45]:
        print("%5.1f%% %5.1f%%  %s:%s  %s" % (100.0 * n / tot, 100.0 * ss / tots, f, ln, src), file=out)


if __name__ == "__main__":
    main()
