# round 2, GPU call 45: the multi-lobe class with its dynamic BSDF / light-sampling routines inlined (libv_dyn)
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 900 python tools/ab_libs.py "libpbrs_gpu.so libv_dyn.so" "c4:0.25 c5:0.125 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab25.log
