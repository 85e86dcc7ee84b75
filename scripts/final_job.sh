python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v12_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency-cases > gpurun_out/launches_v12.log 2>&1
python bench.py --impl reference > gpurun_out/r02_bench_v12_reference_arm.json 2> gpurun_out/r02_bench_v12.err; tail -c 200 gpurun_out/r02_bench_v12_reference_arm.json
python bench.py > gpurun_out/r02_bench_v12.json 2>> gpurun_out/r02_bench_v12.err; tail -c 200 gpurun_out/r02_bench_v12.json
