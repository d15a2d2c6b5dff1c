# round 2, GPU call 14: A/B of the coop slot fill (start mask vs serial) and of deferred unwinds
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1200 python tools/ab_libs.py "libv_base.so libpbrs_gpu.so libv_d1.so libv_d1s2.so libv_d1s4.so libv_d2.so libv_d2v6.so" "c4:0.25 c5:0.125 c3:1.0" 2 2>&1 | tee gpurun_out/r2_ab8.log
