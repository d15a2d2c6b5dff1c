# round 2, GPU call 29: single-lobe classes in registers (libv_lobe0), + diagnostics word in shared memory (libpbrs_gpu), parity first
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libv_r2a.so libv_lobe0.so libpbrs_gpu.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0 c2:1.0" 3 2>&1 | tee gpurun_out/r2_ab16.log
