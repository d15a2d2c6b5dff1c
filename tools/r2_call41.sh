# round 2, GPU call 41: L2 persistence window over the BLAS nodes on the lane streams (carve-out 0 / 32 / 64 MB), C4 at 1/4 frame and C5 at 1/8
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 900 python tools/ab_libs.py "libv_l2.so libv_l2.so@PBRS_L2_PERSIST=32 libv_l2.so@PBRS_L2_PERSIST=64 libv_l2.so@PBRS_L2_PERSIST=100" "c4:0.25 c5:0.125" 2 2>&1 | tee gpurun_out/r2_ab22.log
