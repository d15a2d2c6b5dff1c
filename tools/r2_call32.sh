# round 2, GPU call 32: shade kernels with the queue appends retired one round late (libv_dq) vs at once (libpbrs_gpu)
cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 1500 python tools/ab_libs.py "libpbrs_gpu.so libv_dq.so" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 3 2>&1 | tee gpurun_out/r2_ab18.log
