# round 2, GPU call 25: parity with the aggregated k_generate, default bench, full-size C5 on ONE GPU (the base of the 8-GPU ratio)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --no-cpu > gpurun_out/r2_bench_c4_c.json 2> gpurun_out/r2_bench_c4_c.err; tail -2 gpurun_out/r2_bench_c4_c.err; cat gpurun_out/r2_bench_c4_c.json
python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_bench_c5_full.json 2> gpurun_out/r2_bench_c5_full.err; tail -2 gpurun_out/r2_bench_c5_full.err; cat gpurun_out/r2_bench_c5_full.json
