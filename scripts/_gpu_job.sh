python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_gputest_v12.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v12_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency-cases > gpurun_out/launches_v12.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:^k_ileqg_solve$ -c 1 -f -o gpurun_out/r02_v12_solve python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency-cases > gpurun_out/ncu_v12.log 2>&1; tail -2 gpurun_out/ncu_v12.log
python bench.py > gpurun_out/r02_bench_v12.json 2> gpurun_out/r02_bench_v12.err; tail -c 300 gpurun_out/r02_bench_v12.json
python bench.py --impl reference > gpurun_out/r02_bench_v12_reference_arm.json 2>> gpurun_out/r02_bench_v12.err; tail -c 300 gpurun_out/r02_bench_v12_reference_arm.json
