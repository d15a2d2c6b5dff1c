# round 2, GPU call 17: the same two traversal launches as call 13 (frame 2, batch 4: shadow walk of bounce 0, closest-hit walk of bounce 1) after the L1-pipe changes
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
RUN="python tools/one_frame.py libpbrs_gpu.so c4 0.25 2"
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 131 -c 2 -f -o gpurun_out/r2_p17_trace $RUN > gpurun_out/r2_p17_ncu1.log 2>&1; tail -2 gpurun_out/r2_p17_ncu1.log
