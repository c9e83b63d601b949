"""Summarise an .ncu-rep: per-kernel headline metrics + per-opcode executed instruction counts (source page).
usage: python tools/ncu_summary.py report.ncu-rep [units_per_launch_for_kernel_regex=N ...]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
units = dict(a.split("=") for a in sys.argv[2:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, un = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__warps_eligible.avg.per_cycle_active"]
seen = set()
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    print("=== %s (id %s)" % (name[:100], r[ix["ID"]]))
    for w in want:
        if w in ix:
            print("  %-85s %s %s" % (w, r[ix[w]], un[ix[w]]))
    short = re.sub(r"\(.*", "", name)
    if short in seen:
        continue
    seen.add(short)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(int(r[ix["ID"]])),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    h2 = None
    byop, samples, tot = collections.Counter(), collections.Counter(), 0
    for sr in srows:
        if "Instructions Executed" in sr and "Source" in sr:
            h2 = {h: i for i, h in enumerate(sr)}
            continue
        if h2 is None or len(sr) < len(h2):
            continue
        try:
            n = int(sr[h2["Instructions Executed"]]); smp = int(sr[h2["# Samples"]])
        except ValueError:
            continue
        toks = sr[h2["Source"]].strip().split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += n; samples[op] += smp; tot += n
    u = None
    for k, v in units.items():
        if re.search(k, name):
            u = float(v)
    print("  -- executed warp instructions: %d%s" % (tot, " = %.1f per unit" % (tot / u) if u else ""))
    for op, n in byop.most_common(22):
        print("     %-10s %12d %s  stall-samples %d" % (op, n, "%8.2f/unit" % (n / u) if u else "", samples[op]))
