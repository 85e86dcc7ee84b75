# final-tree evidence of a round: smoke, GPU tests, launch list, ncu --set full of the throughput kernel, both bench arms
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_gputest_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency-cases > gpurun_out/launches_final.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:^k_ileqg_solve$ -c 1 -f -o gpurun_out/r02_final_solve python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency-cases > gpurun_out/ncu_final.log 2>&1; tail -1 gpurun_out/ncu_final.log
python bench.py --impl reference > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final_reference_arm.json
python bench.py > gpurun_out/r02_bench_final.json 2>> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final.json
