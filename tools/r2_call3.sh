# round 2, GPU call 3: parity (incl. the new film paths) + A/B of the cooperative closest-hit leaf phase and refill / vote knobs
export PBRS_WRITE_OUTLIERS=1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libold.so libpbrs_gpu.so libv_nocoop.so libv_local.so libv_idle12.so libv_idle8.so libv_idle8l.so libv_b7.so libpbrs_gpu.so@PBRS_LEAF_VOTE_CLOSEST=4 libpbrs_gpu.so@PBRS_LEAF_VOTE_CLOSEST=6 libpbrs_gpu.so@PBRS_LEAF_VOTE_CLOSEST=12 libpbrs_gpu.so@PBRS_LEAF_VOTE_CLOSEST=16" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab2.log
