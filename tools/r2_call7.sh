# round 2, GPU call 7: rolled step loop (code size) -- steps 2 / 3 / 4, cooperative leaf phase on / off
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libold.so libpbrs_gpu.so libv_steps2.so libv_steps4.so libv_nocoop.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab6.log
