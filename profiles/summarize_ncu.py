"""Turn an .ncu-rep (one kernel, `ncu --set full --import-source on`) into the markdown summaries kept
under profiles/:  python profiles/summarize_ncu.py report.ncu-rep "title" "command" > summary.md"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
           "dram__bytes_write.sum"] + [
    "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k for k in
    ("wait", "barrier", "not_selected", "branch_resolving", "short_scoreboard", "long_scoreboard",
     "no_instruction", "math_pipe_throttle")]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep, title, command = sys.argv[1:4]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    h, u, v = rows[0], rows[1], rows[2]
    print("# %s\n\nCommand (1 x B200): `%s` (after the same command exited 0 without ncu)\n" % (title, command))
    print("| metric | unit | value |\n|---|---|---|")
    for m in METRICS:
        if m in h:
            i = h.index(m)
            print("| %s | %s | %s |" % (m, u[i], v[i]))
    cur = hdr = None
    lines = defaultdict(lambda: [0, 0, ""])
    byfile = defaultdict(int)
    for r in csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if cur and hdr and len(r) > 10 and r[0].isdigit():
            try:
                inst, smp = int(r[ii]), int(r[si])
            except ValueError:
                continue
            key = (cur, int(r[0]))
            lines[key][0] += inst
            lines[key][1] += smp
            lines[key][2] = r[1].strip()[:90]
            byfile[cur] += inst
    tot = sum(x[0] for x in lines.values()) or 1
    stot = sum(x[1] for x in lines.values()) or 1
    print("\n## Executed warp-instructions / stall samples per CUDA source line (top 25)\n\n```")
    print("total warp-inst", tot, "samples", stot)
    for k, x in sorted(byfile.items(), key=lambda kv: -kv[1])[:8]:
        print("%-28s %12d %5.1f%%" % (k, x, 100.0 * x / tot))
    print()
    for k, x in sorted(lines.items(), key=lambda kv: -kv[1][0])[:25]:
        print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * x[0] / tot, 100.0 * x[1] / stot, k[0], k[1], x[2]))
    print("```")


if __name__ == "__main__":
    main()
