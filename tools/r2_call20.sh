# round 2, GPU call 20: full GPU suite + default bench at HEAD (256-bit loads, packed stack, deferred any-hit unwinds)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py > gpurun_out/r2_bench_c4_b.json 2> gpurun_out/r2_bench_c4_b.err; tail -3 gpurun_out/r2_bench_c4_b.err; cat gpurun_out/r2_bench_c4_b.json
