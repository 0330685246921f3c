"""Summarises an .ncu-rep (read here, no GPU needed): key raw metrics of the first profiled launch,
opcode histogram weighted by executed instructions, and the hottest SASS lines by stall samples.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
       "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_uniform.sum",
       "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
       "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_average_branch_targets_threads_uniform.pct",
       "sm__sass_thread_inst_executed_op_fp32_pred_on.sum", "smsp__inst_executed_op_branch.sum"]

out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H, U = rows[0], rows[1]
for k, r in enumerate(rows[2:4]):
    print("== launch %d: %s" % (k, r[H.index("Kernel Name")][:100]))
    for i, h in enumerate(H):
        if h in RAW or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            print("  %-88s %s %s" % (h, r[i], U[i]))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
H = rows[hdr[0]]
sec = rows[hdr[0] + 1:(hdr[1] - 1 if len(hdr) > 1 else len(rows))]
ie, src, smp = H.index("Instructions Executed"), H.index("Source"), H.index("# Samples")
tot = 0
ops = collections.Counter()
lines = []
for r in sec:
    try:
        n = int(r[ie])
    except (ValueError, IndexError):
        continue
    s = r[src].strip()
    toks = s.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    ops[op.split(".")[0]] += n
    tot += n
    lines.append((int(r[smp] or 0), n, s))
print("== SASS of the first launch: %d static instructions, %d executed warp-instructions" % (len(lines), tot))
for k, v in ops.most_common(28):
    print("  %-10s %12d  %5.1f %%" % (k, v, 100.0 * v / tot))
print("== hottest SASS lines by stall samples")
ts = sum(l[0] for l in lines)
for s_, n, txt in sorted(lines, reverse=True)[:25]:
    print("  %5.1f %%  exec %10d  %s" % (100.0 * s_ / max(ts, 1), n, txt))
