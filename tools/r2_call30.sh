# round 2, GPU call 30: split vs one-piece shade kernels again after the local-memory work; ncu of k_surface + k_scatter<Lambert> (bounce 0, batch 4)
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
timeout 900 python tools/ab_libs.py "libpbrs_gpu.so@PBRS_SHADE_SPLIT=0 libpbrs_gpu.so@PBRS_SHADE_SPLIT=1" "c4:0.25 c5:0.125 c3:1.0 c1:1.0" 3 2>&1 | tee gpurun_out/r2_ab17.log
RUN="python tools/one_frame.py libpbrs_gpu.so c4 0.25 2"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_surface|k_scatter|k_shade" --csv --log-file gpurun_out/r2_p30_shade_list.csv $RUN > gpurun_out/r2_p30_l.log 2>&1
S=$(python tools/ncu_pick.py gpurun_out/r2_p30_shade_list.csv "k_surface" 16 1000); echo shade skip $S
ncu --set full --clock-control none --import-source on -k regex:"k_surface|k_scatter|k_shade" -s $S -c 2 -f -o gpurun_out/r2_p30_shade $RUN > gpurun_out/r2_p30_ncu.log 2>&1; tail -2 gpurun_out/r2_p30_ncu.log
