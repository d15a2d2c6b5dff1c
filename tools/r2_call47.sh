# round 2, GPU call 47: texture_value and eval_env inlined (libv_te)
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 900 python tools/ab_libs.py "libpbrs_gpu.so libv_te.so" "c4:0.25 c5:0.125 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab26.log
