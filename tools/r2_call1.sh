# round 2, GPU call 1: parity of the walker rewrite + A/B of its variants against the round-1 library
export PBRS_WRITE_OUTLIERS=1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
cp profiles/parity_outliers.json gpurun_out/ 2>/dev/null
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libold.so libpbrs_gpu.so libv_local.so libv_steps1.so libv_fma.so libv_b6.so libv_b7.so libv_idle12.so libv_s16.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab1.log
