# round 2, GPU call 27: shade kernels with reconstruct_hit / the area-light sampling helpers inlined (struct arguments out of local memory)
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libpbrs_gpu.so libv_ir.so libv_ia.so libv_ira.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 3 2>&1 | tee gpurun_out/r2_ab14.log
