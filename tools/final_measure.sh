# Round-end evidence for the committed kernels (one B200, through gpurun; leaves only small files in gpurun_out/):
#   1. GPU suite with PBRS_WRITE_OUTLIERS=1 (measured outlier counts), smoke()
#   2. the default bench command + its reference arm; bench lines of the other BASELINE configs
#   3. ncu metric pass over all bounces of two whole mid-frame batches of the full-size C4 frame -> profiles/r2_traffic.json
#   4. ncu launch list (gpu__time_duration only) of every kernel of a quarter-size C4 frame
#   5. ncu --set full captures -> markdown summaries (tools/ncu_md.py) of: any-hit walk of bounce 0 + closest-hit walk of bounce 1 (C4),
#      the Lambert shade kernel of bounce 0 (C4), the closest-hit walk of bounce 1 (C5)
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
O=gpurun_out
PBRS_WRITE_OUTLIERS=1 timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
cp profiles/parity_outliers.json profiles/parity_outliers_samples.json $O/ 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > $O/final_bench_c4.json 2> $O/final_bench_c4.err; tail -2 $O/final_bench_c4.err; cat $O/final_bench_c4.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_c4_reference_arm.json 2> $O/final_ref.err; cat $O/final_bench_c4_reference_arm.json
for w in c1 c2 c3; do python bench.py --workload $w --steps 10 --warmup 3 > $O/final_bench_$w.json 2> $O/final_bench_$w.err; done
python bench.py --workload c5 --frame-scale 0.125 --steps 2 --warmup 3 > $O/final_bench_c5_eighth.json 2> $O/final_bench_c5e.err
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__sass_inst_executed_op_shared.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
# full-size frame: 128 batches x 10 traversal launches; frame 2 starts at 1280, batch 64 at 1920
timeout 900 ncu --metrics $METRICS --clock-control none -k regex:k_trace -s 1920 -c 20 -f -o /tmp/r2_trace_metrics python tools/one_frame.py libpbrs_gpu.so c4 1.0 2 > $O/final_ncu_metrics.log 2>&1; tail -1 $O/final_ncu_metrics.log
python tools/ncu_traffic.py /tmp/r2_trace_metrics.ncu-rep c4 --store 2>&1 | tail -3
cp profiles/r2_traffic.json $O/
# launch list: every kernel of a quarter-size frame (same batch size, same launches per batch as the full frame).  NOT the full bench
# command: ncu intercepts every launch it skips, and the 25 000 launches before the timed frame took more than 25 minutes when tried.
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/final_launch_list_c4_quarter.csv python tools/one_frame.py libpbrs_gpu.so c4 0.25 2 > $O/final_ncu_list.log 2>&1; tail -1 $O/final_ncu_list.log | cut -c1-200
# source-level captures -> markdown
mkdir -p /tmp/cub && (cd /tmp/cub && rm -f *.cubin *.dis && cuobjdump -xelf all $GRAFT_REPO_ROOT/pbrs_b200/lib/libpbrs_gpu.so > /dev/null && nvdisasm -g -c kernels.sm_100a.cubin > kernels.sm_100a.cubin.dis 2>/dev/null)
RUN="python tools/one_frame.py libpbrs_gpu.so c4 0.25 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 131 -c 2 -f -o /tmp/r2_trace_c4 $RUN > $O/final_ncu_a.log 2>&1
python tools/ncu_md.py /tmp/r2_trace_c4.ncu-rep "k_trace at HEAD: C4 at 1/4 frame (same 16 Mi-path batches as the full frame), frame 2, batch 4 -- any-hit walk of bounce 0, closest-hit walk of bounce 1" --lines k_traceILb1ELb0ELb0:1 k_traceILb0ELb0ELb0:2 > $O/final_ncu_trace_c4.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 327 -c 1 -f -o /tmp/r2_shade_c4 $RUN > $O/final_ncu_b.log 2>&1
python tools/ncu_md.py /tmp/r2_shade_c4.ncu-rep "k_shade<Lambert, path> at HEAD: C4 at 1/4 frame, frame 2, batch 4, bounce 0 (16.7 M paths)" --lines k_shadeILi2ELi1E:0 > $O/final_ncu_shade_c4.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 142 -c 1 -f -o /tmp/r2_trace_c5 python tools/one_frame.py libpbrs_gpu.so c5 0.125 2 > $O/final_ncu_c.log 2>&1
python tools/ncu_md.py /tmp/r2_trace_c5.ncu-rep "k_trace at HEAD on C5 (10 000 instances, 1/8 frame), frame 2, batch 4: closest-hit walk of bounce 1" --lines k_traceILb0ELb0ELb0:0 > $O/final_ncu_trace_c5.md
ls -la $O
