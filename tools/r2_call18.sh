# round 2, GPU call 18: L1 carve-out of the traversal kernels; shared-memory ring stack re-measured on top of the L1-pipe changes
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libpbrs_gpu.so libv_carve.so libv_carve6.so libv_ring4.so libv_ring8.so libv_ring16.so" "c4:0.25 c5:0.125 c3:1.0" 2 2>&1 | tee gpurun_out/r2_ab11.log
