# round 2, GPU call 38: fresnel_eval inlined (the Lobe of the microfacet / specular classes out of local memory)
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 900 python tools/ab_libs.py "libpbrs_gpu.so libv_fr.so" "c4:0.25 c5:0.125 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab21.log
