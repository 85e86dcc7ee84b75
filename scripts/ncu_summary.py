"""Summarise one kernel of an ncu report into the text files committed under profiles/ (the .ncu-rep itself is scratch):
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep "<command that was profiled>" "<one-line description>" > profiles/x.txt
Prints the launch shape, duration, pipe / issue utilisation, DRAM traffic, the stall breakdown (per issue) and, from the
source page, the executed-instruction mix by opcode with each opcode's share of the stall samples."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, cmd, desc = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, v = rows[0], rows[1], rows[2]
print(cmd)
print("kernel:", v[h.index("Kernel Name")], "--", desc)
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for k in want:
    if k in h:
        print(f"{k} [{units[h.index(k)]}] = {v[h.index(k)]}")
stalls = [(float(v[i]), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for i, k in enumerate(h)
          if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
print("stall cycles per issued instruction:", ", ".join(f"{n} {x:.2f}" for x, n in sorted(stalls, reverse=True) if x >= 0.01))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh = srows[1]
iS, iE, iM = sh.index("Source"), sh.index("Instructions Executed"), sh.index("# Samples")
ops, smp = Counter(), Counter()
for r in srows[2:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(r[iE]); smp[op] += int(r[iM])
tot, tots = sum(ops.values()), max(sum(smp.values()), 1)
print("instruction mix (warp instructions; share of stall samples):",
      ", ".join(f"{o} {100 * c / tot:.1f}% ({100 * smp[o] / tots:.1f}%)" for o, c in ops.most_common(14)))
