"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + executed-instruction mix by opcode.
usage: python scripts/ncu_summary.py gpurun_out/dense_v2.ncu-rep [out.md]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
STALL = "smsp__average_warps_issue_stalled_"
print("## %s\n" % rep, file=out)
print("| metric | unit | value |\n|---|---|---|", file=out)
for h, u, v in zip(hdr, units, vals):
    if h in KEYS or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
        print("| %s | %s | %s |" % (h.replace(STALL, "stall:").replace("_per_issue_active.ratio", ""), u, v), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
iS, iE, iSm = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
byop, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= iE:
        continue
    m = re.match(r"\s*(@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", r[iS])
    op = m.group(2) if m else r[iS]
    n = int(r[iE] or 0)
    byop[op] += n; samp[op] += int(r[iSm] or 0); tot += n
print("\nExecuted warp-instructions: %d\n\n| opcode | warp-inst | share | stall samples |\n|---|---|---|---|" % tot, file=out)
for op, n in byop.most_common(24):
    print("| %s | %d | %.1f%% | %d |" % (op, n, 100.0 * n / tot, samp[op]), file=out)
