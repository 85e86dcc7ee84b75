python -m pytest tests/test_gpu_parity.py tests/test_edge_cases.py tests/test_fleet_scheduling.py -m gpu -x -q 2>&1 | tail -2
for v in libratilqr_b200.so libvariant_depth1.so libratilqr_b200.so libvariant_depth1.so; do
RATILQR_B200_LIB=$PWD/ratilqr.jl_b200/csrc/$v python bench.py --no-cpu-baseline > gpurun_out/ab_$v.log 2>/dev/null
python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open(f"gpurun_out/ab_{v}.log").read().strip().splitlines()[-1])
print(v, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "mpc", round(d["mpc_step"]["ms_per_fleet_step"],1), "c2_single", round(d["c2_single"]["ms_per_batch"],2), d["clocks"]["sm_mhz"])
PY
done
