# round 2, GPU call 16: any-hit deferred unwinds + packed park words (main) vs tuning variants on top of the 256-bit loads
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libv_base.so libpbrs_gpu.so libv_nodefer.so libv_fma.so libv_b9.so libv_b10.so libv_va8.so libv_va16.so libv_ri4.so libv_ri12.so" "c4:0.25 c5:0.125 c3:1.0" 2 2>&1 | tee gpurun_out/r2_ab10.log
