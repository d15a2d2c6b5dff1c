# round 2, GPU call 13: source-level captures picked from the launch list of call 12 (frame 2, batch 4):
# shadow walk of bounce 0 + closest-hit walk of bounce 1; k_surface + k_scatter<Lambert> of bounce 0
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
RUN="python tools/one_frame.py libpbrs_gpu.so c4 0.25 2"
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 131 -c 2 -f -o gpurun_out/r2_p13_trace $RUN > gpurun_out/r2_p13_ncu1.log 2>&1; tail -2 gpurun_out/r2_p13_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:"k_surface|k_scatter|k_shade" -s 457 -c 2 -f -o gpurun_out/r2_p13_shade $RUN > gpurun_out/r2_p13_ncu2.log 2>&1; tail -2 gpurun_out/r2_p13_ncu2.log
ls -la gpurun_out
