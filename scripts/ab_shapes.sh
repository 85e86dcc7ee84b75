#!/bin/bash
# Same-box A/B of launch shapes of the throughput kernel (RATILQR_SOLVE_SHAPE) in a -DRL_TUNE_SHAPES build:
#   scripts/ab_shapes.sh <out-file> <rounds> <lib> <shape> <shape> ...      ("d" = library default)
out=$1; rounds=$2; lib=$3; shift 3
: > $out
for r in $(seq $rounds); do
for sh in "$@"; do
  if [ "$sh" = d ]; then unset RATILQR_SOLVE_SHAPE; else export RATILQR_SOLVE_SHAPE=$sh; fi
  RATILQR_DEBUG=1 RATILQR_B200_LIB=$PWD/$lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-latency-cases > gpurun_out/ab_tmp.log 2>gpurun_out/ab_tmp.err || { echo "shape $sh FAILED" >> $out; tail -3 gpurun_out/ab_tmp.err >> $out; continue; }
  grep -m1 "k_ileqg_solve_r" gpurun_out/ab_tmp.err >> $out
  python - $sh >> $out <<'PY'
import json,sys
d=json.loads(open("gpurun_out/ab_tmp.log").read().strip().splitlines()[-1])
print("shape", sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), "mhz", d["clocks"]["sm_mhz"], "converged", d["converged_instances"])
PY
done
done
unset RATILQR_SOLVE_SHAPE
cat $out
