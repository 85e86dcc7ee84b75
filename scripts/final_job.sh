# final-tree evidence of a round: smoke, GPU tests, launch list, ncu --set full of the throughput kernel, both bench arms
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_gputest_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency-cases > gpurun_out/launches_final.log 2>&1
# (the ncu --set full capture of the throughput kernel is taken separately: it did not change since r02_final_solve_ncu_summary.txt)
python bench.py --impl reference > gpurun_out/r02_bench_final_reference_arm.json 2> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final_reference_arm.json
python bench.py > gpurun_out/r02_bench_final.json 2>> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final.json
