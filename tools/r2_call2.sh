# round 2, GPU call 2: ncu --set full of the first closest-hit / any-hit launches, round-1 library vs the walker rewrite
for lib in libold.so libpbrs_gpu.so libv_local.so; do
  PYTHONPATH=. python tools/one_frame.py $lib c4 0.25 > gpurun_out/plain_$lib.log 2>&1 && \
  PYTHONPATH=. ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -f -o gpurun_out/r2_trace_${lib%.so} python tools/one_frame.py $lib c4 0.25 > gpurun_out/ncu_$lib.log 2>&1
  tail -2 gpurun_out/plain_$lib.log gpurun_out/ncu_$lib.log
done
