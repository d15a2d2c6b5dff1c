# round 2, GPU call 33: traversal without the per-push overflow check (commit bounds the depth); closest-hit leaf vote re-checked
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libpbrs_gpu.so libv_nochk.so libpbrs_gpu.so@PBRS_LEAF_VOTE_CLOSEST=6 libpbrs_gpu.so@PBRS_LEAF_VOTE_CLOSEST=10" "c4:0.25 c5:0.125 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab19.log
