# round 2, GPU call 2b: ncu --set full of mid-frame closest-hit / any-hit launches (batch 4 of 8, bounces 0 and 1), round-1 library vs the walker rewrite
for lib in libold.so libpbrs_gpu.so libv_local.so; do
  PYTHONPATH=. python tools/one_frame.py $lib c4 0.25 > gpurun_out/plain_$lib.log 2>&1 && \
  PYTHONPATH=. ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 4 -f -o gpurun_out/r2_trace_${lib%.so} python tools/one_frame.py $lib c4 0.25 > gpurun_out/ncu_$lib.log 2>&1
  tail -n 2 gpurun_out/plain_$lib.log gpurun_out/ncu_$lib.log
done
