# round 2, GPU call 22: bench lines of the other BASELINE configs at HEAD (C1, C2, C3 full size; C5 at 1/8 frame), CPU baseline beside each
cd $GRAFT_REPO_ROOT
for w in c1 c2 c3; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_bench_${w}_b.json 2> gpurun_out/r2_bench_${w}_b.err; tail -2 gpurun_out/r2_bench_${w}_b.err; cat gpurun_out/r2_bench_${w}_b.json; done
python bench.py --workload c5 --frame-scale 0.125 --steps 2 --warmup 3 > gpurun_out/r2_bench_c5e_b.json 2> gpurun_out/r2_bench_c5e_b.err; tail -2 gpurun_out/r2_bench_c5e_b.err; cat gpurun_out/r2_bench_c5e_b.json
