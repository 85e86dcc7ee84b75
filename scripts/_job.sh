python -m pytest tests/test_spec_kernel.py tests/test_gpu_parity.py tests/test_baseline_sizes.py tests/test_literal_pin.py tests/test_edge_cases.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_trig.json 2> gpurun_out/bench_trig.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_trig.json').read().strip().splitlines()[-1])
print("c3 roofline", d["roofline_c3"]["ms_per_launch"], d["roofline_c3"]["solves_per_sec"], "lat", [d["latency"][k]["gpu_ms"] for k in d["latency"] if k!="cpu_threads"])
PY
