# round 2, GPU call 43: flat-box tier of the slow box test (libpbrs_gpu) vs before (libv_r2b); parity first
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
PYTHONPATH=. timeout 900 python tools/ab_libs.py "libv_r2b.so libpbrs_gpu.so" "c1:1.0 c2:1.0 c5:0.125 c4:0.25 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab24.log
