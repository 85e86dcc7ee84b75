#!/bin/bash
# Same-box A/B of library variants: scripts/ab_variants.sh <out-file> <rounds> <lib1> <lib2> ...
# (library paths relative to the repo root; every variant runs `rounds` times, interleaved)
out=$1; rounds=$2; shift 2
: > $out
for r in $(seq $rounds); do
for v in "$@"; do
  RATILQR_B200_LIB=$PWD/$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_tmp.log 2>gpurun_out/ab_tmp.err || { echo "$v FAILED" >> $out; tail -3 gpurun_out/ab_tmp.err >> $out; continue; }
  python - $v >> $out <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open("gpurun_out/ab_tmp.log").read().strip().splitlines()[-1])
lat=d.get("latency",{})
print(v, round(d["value"]), round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), "mhz", d["clocks"]["sm_mhz"], "c2_single", round(d["c2_single"]["ms_per_batch"],3), "e2e", round(d["e2e"]["value"]), "mpc_fleet", round(d["mpc_step"]["ms_per_fleet_step"],1), "lat(c2_1x1024, mpc_single, c3_nm ms)", [round(lat[k]["gpu_ms"],2) for k in ("c2_single_1x1024","mpc_step_single_problem","c3_rat_ilqr_pp_quadrotor") if k in lat])
PY
done
done
cat $out
