# Round-end evidence for the default bench command (C4, full size): plain run, ncu launch list, one ncu --set full capture.
set -x
python bench.py > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 7896 -c 420 --csv --log-file gpurun_out/final_launches.csv python bench.py --no-cpu > gpurun_out/final_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1880 -c 2 -f -o gpurun_out/final_trace_c4 python bench.py --no-cpu > gpurun_out/final_ncu2.log 2>&1
tail -2 gpurun_out/final_ncu1.log gpurun_out/final_ncu2.log
cat gpurun_out/final_plain.json
