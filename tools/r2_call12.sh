# round 2, GPU call 12: launch list of a 2-frame run of C4 at 1/4 frame, then source-level captures picked from it
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
RUN="python tools/one_frame.py libpbrs_gpu.so c4 0.25 2"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r2_p12_trace_list.csv $RUN > gpurun_out/r2_p12_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_surface|k_scatter|k_shade" --csv --log-file gpurun_out/r2_p12_shade_list.csv $RUN > gpurun_out/r2_p12_l2.log 2>&1
# frame 2 starts after 90 traversal launches; the 12th heavy closest-hit launch after that is batch 2 / bounce 1
S=$(python tools/ncu_pick.py gpurun_out/r2_p12_trace_list.csv "k_trace<0" 57 1000); echo trace skip $S
ncu --set full --clock-control none --import-source on -k regex:k_trace -s $S -c 2 -f -o gpurun_out/r2_p12_trace $RUN > gpurun_out/r2_p12_ncu1.log 2>&1; tail -2 gpurun_out/r2_p12_ncu1.log
S=$(python tools/ncu_pick.py gpurun_out/r2_p12_shade_list.csv "k_surface" 57 200); echo shade skip $S
ncu --set full --clock-control none --import-source on -k regex:"k_surface|k_scatter|k_shade" -s $S -c 2 -f -o gpurun_out/r2_p12_shade $RUN > gpurun_out/r2_p12_ncu2.log 2>&1; tail -2 gpurun_out/r2_p12_ncu2.log
ls -la gpurun_out
