# Round-end evidence for the default bench command (C4, full size): plain run, ncu launch list, ncu metric
# pass over two whole batches of traversal launches (all bounces), one ncu --set full capture with source
# of a mid-frame bounce-0 / bounce-1 pair, and one of the shade kernels of a stage.
# (launch arithmetic: 20 frames of 10752 launches precede nothing -- ncu counts from process start, the
# untimed counter frame comes first: its 640 batches x 16.8 launches; -s values below land mid-frame 2)
set -x
python bench.py > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err
# the ncu passes profile a shorter run of the same workload (frames: counters, 1 warm-up, 1 timed, 1 staged, e2e)
CMD="python bench.py --no-cpu --steps 1 --warmup 1"
$CMD > gpurun_out/final_plain_short.json 2> gpurun_out/final_plain_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 16000 -c 420 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu1.log 2>&1
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__sass_inst_executed_op_shared.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
ncu --metrics $METRICS --clock-control none -k regex:k_trace -s 3000 -c 20 -f -o gpurun_out/final_trace_metrics $CMD > gpurun_out/final_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3000 -c 4 -f -o gpurun_out/final_trace_c4 $CMD > gpurun_out/final_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_surface|k_scatter" -s 1785 -c 14 -f -o gpurun_out/final_shade_c4 $CMD > gpurun_out/final_ncu4.log 2>&1
tail -n 2 gpurun_out/final_ncu1.log gpurun_out/final_ncu2.log gpurun_out/final_ncu3.log gpurun_out/final_ncu4.log
cat gpurun_out/final_plain.json
