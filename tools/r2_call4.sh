# round 2, GPU call 4: ray sorting between bounces (on / off), parity first
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libpbrs_gpu.so@PBRS_SORT_RAYS=0 libpbrs_gpu.so@PBRS_SORT_RAYS=1" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab3.log
