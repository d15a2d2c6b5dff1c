# round 2, GPU call 23: class-mask launches + in-place one-leaf meshes (libpbrs_gpu) vs the build before (libv_r2a), parity first
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libv_r2a.so libpbrs_gpu.so" "c1:1.0 c2:1.0 c3:1.0 c4:0.25 c5:0.125" 3 2>&1 | tee gpurun_out/r2_ab13.log
