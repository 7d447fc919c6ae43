"""Summarise an .ncu-rep (read on the CPU box): per captured launch, the metrics DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--sass K] > profiles/xxx.txt
--sass K additionally prints the hottest SASS lines (stall samples) of captured launch K.
"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
sass_k = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_bytes.sum", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
print(f"# {rep}: {len(data)} captured launches (ncu --set full --clock-control none)")
for k, r in enumerate(data):
    print(f"\n## launch {k}: {r[hdr.index('Kernel Name')][:100]}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:95s} {r[i]:>16s} {units[i]}")
if sass_k is not None:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(sass_k), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    iS, iI, isrc, ia = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source"), h.index("Avg. Threads Executed")
    d = [r for r in rows[2:] if len(r) > iI and r[iS].isdigit()]
    half = len(d) // 2 if len(d) > 2 and d[0][isrc] == d[len(d) // 2][isrc] else len(d)   # the page lists the kernel twice
    d = d[:half]
    tot = sum(int(r[iS]) for r in d); toti = sum(int(r[iI]) for r in d)
    print(f"\n## SASS of launch {sass_k}: {len(d)} instructions, {toti} warp-instructions executed, {tot} stall samples; lines with >= 0.7 % of the samples")
    print("   idx  samples  warp-inst  thr  sass")
    for k, r in enumerate(d):
        if int(r[iS]) >= 0.007 * tot:
            print(f"  {k:4d} {int(r[iS]):7d} {int(r[iI]):10d} {r[ia]:>4s}  {r[isrc].strip()[:100]}")
    # instruction mix by opcode
    mix = {}
    for r in d:
        op = r[isrc].strip().split()
        op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
        mix[op] = mix.get(op, 0) + int(r[iI])
    print("   opcode mix (warp-instructions): " + ", ".join(f"{k}={v}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:18]))
