# round 2, GPU call 15: 256-bit record loads + packed (link, t_low) stack entries, with / without deferred unwinds
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1200 python tools/ab_libs.py "libv_base.so libv_ld128.so libpbrs_gpu.so libv_d1.so libv_d1s2.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 2 2>&1 | tee gpurun_out/r2_ab9.log
