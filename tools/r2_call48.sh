# round 2, GPU call 48: one-piece shade kernels at 5 blocks per SM (96 registers) after the local-memory work
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 900 python tools/ab_libs.py "libpbrs_gpu.so libv_sh5.so" "c4:0.25 c5:0.125 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab27.log
