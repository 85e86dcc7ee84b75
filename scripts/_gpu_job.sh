python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sincos_any.json 2> gpurun_out/bench_sincos_any.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_sincos_any.json').read().strip().splitlines()[-1])
print("value", d["value"], "c4 ms", d["roofline_c4"]["ms_per_solve"], d["roofline_c4"]["rollouts_per_sec"], "c3", d["roofline_c3"]["ms_per_launch"], d["roofline_c3"]["solves_per_sec"], "mc_eval", d["mpc_step"].get("mc_eval"), "lat", [d["latency"][k]["gpu_ms"] for k in d["latency"] if k!="cpu_threads"])
PY
