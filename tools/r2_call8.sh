# round 2, GPU call 8: split shade kernels (surface + scatter) vs the one-piece kernels, parity first
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libold.so libv_mono.so libpbrs_gpu.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab7.log
