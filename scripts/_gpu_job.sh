python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_v12.json 2> gpurun_out/r02_bench_v12.err; tail -c 600 gpurun_out/r02_bench_v12.json
