python bench.py --no-cpu-baseline > gpurun_out/bench_v9.log 2> gpurun_out/bench_v9.err
python bench.py --no-cpu-baseline --no-profile-warm > gpurun_out/bench_v9_noprofile.log 2>> gpurun_out/bench_v9.err
RATILQR_PROFILE_ORDER=0 python bench.py --no-cpu-baseline > gpurun_out/bench_v9_profileoff.log 2>> gpurun_out/bench_v9.err
for f in bench_v9 bench_v9_noprofile bench_v9_profileoff; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
d=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
print(f, round(d["value"]), round(d["ms_per_step"],1), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), round(d["mpc_step"]["ms_per_fleet_step"],1), d["clocks"])
PY
done
tail -3 gpurun_out/bench_v9.err
