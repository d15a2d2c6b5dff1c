# round 2, GPU call 21: C5 (1/8 frame) launch list + source-level capture of a mid-frame closest-hit walk of bounce 1 and the shadow walk after it
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
RUN="python tools/one_frame.py libpbrs_gpu.so c5 0.125 2"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r2_p21_trace_list.csv $RUN > gpurun_out/r2_p21_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 142 -c 2 -f -o gpurun_out/r2_p21_trace_c5 $RUN > gpurun_out/r2_p21_ncu1.log 2>&1; tail -2 gpurun_out/r2_p21_ncu1.log
