# Round-2 evidence that feeds committed files (run through gpurun; everything it leaves in gpurun_out/ is small):
#  1. the GPU suite with PBRS_WRITE_OUTLIERS=1 -> measured outlier counts (tests/golden/parity_outliers*.json are pinned from them)
#  2. ncu metric pass over all five bounces of two whole mid-frame batches of the full-size C4 frame -> profiles/r2_traffic.json
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
export PBRS_WRITE_OUTLIERS=1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
unset PBRS_WRITE_OUTLIERS
cp profiles/parity_outliers.json profiles/parity_outliers_samples.json gpurun_out/ 2>/dev/null
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__sass_inst_executed_op_shared.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
# full-size frame: 128 batches x 10 traversal launches; frame 2 starts at 1280, batch 64 at 1920
timeout 1500 ncu --metrics $METRICS --clock-control none -k regex:k_trace -s 1920 -c 20 -f -o /tmp/r2_trace_metrics python tools/one_frame.py libpbrs_gpu.so c4 1.0 2 > gpurun_out/r2_ev_ncu.log 2>&1; tail -2 gpurun_out/r2_ev_ncu.log
python tools/ncu_traffic.py /tmp/r2_trace_metrics.ncu-rep c4 --store 2>&1 | tail -4
cp profiles/r2_traffic.json gpurun_out/
