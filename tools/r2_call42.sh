# round 2, GPU call 42: path-state loads / stores marked evict-first (ld/st.global.cs) so that the streaming state does not push the BVH out of L2
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 900 python tools/ab_libs.py "libpbrs_gpu.so libv_cs.so" "c4:0.25 c5:0.125 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab23.log
