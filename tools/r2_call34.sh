# round 2, GPU call 34: three batch lanes vs two; ncu launch list (all kernels) of a quarter-size C4 frame at HEAD
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
timeout 900 python tools/ab_libs.py "libpbrs_gpu.so libv_l3.so" "c4:0.25 c5:0.125 c3:1.0" 3 2>&1 | tee gpurun_out/r2_ab20.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_list_head.csv python tools/one_frame.py libpbrs_gpu.so c4 0.25 2 > gpurun_out/r2_p34.log 2>&1; tail -1 gpurun_out/r2_p34.log
