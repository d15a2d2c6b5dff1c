# round 2, GPU call 28: + __grid_constant__ kernel parameters (libpbrs_gpu) vs the same without (libv_nogc) vs inlining only as measured before (libv_ira) vs before (libv_r2a)
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libv_r2a.so libv_ira.so libv_nogc.so libpbrs_gpu.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 3 2>&1 | tee gpurun_out/r2_ab15.log
