# round 2, GPU call 19: occupancy variants of the split shade kernels (blocks per SM of k_scatter, k_surface, k_shade)
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libpbrs_gpu.so libv_sc5.so libv_sc6.so libv_su8.so libv_su5.so libv_sh5.so" "c4:0.25 c5:0.125 c3:1.0" 2 2>&1 | tee gpurun_out/r2_ab12.log
