# round 2, GPU call 5: parity + A/B of stack placement / prefetch / refill knobs on the trimmed walker
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libold.so libpbrs_gpu.so libv_local.so libv_s4.so libv_pf.so libv_steps3.so libv_idle5.so libv_chunk128.so libv_anyvote8.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab4.log
