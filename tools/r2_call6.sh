# round 2, GPU call 6: parity + A/B of the phased shade kernels, then the full-size default bench with its e2e leg
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libold.so libpbrs_gpu.so libv_ph128.so libv_ph256.so libv_ph512.so libv_sb5.so libv_smem.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab5.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_c4_a.json 2> gpurun_out/r2_bench_c4_a.err; tail -3 gpurun_out/r2_bench_c4_a.err; cat gpurun_out/r2_bench_c4_a.json
python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c2_a.json 2> gpurun_out/r2_bench_c2_a.err; tail -3 gpurun_out/r2_bench_c2_a.err; cat gpurun_out/r2_bench_c2_a.json
