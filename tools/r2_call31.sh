# round 2, GPU call 31: parity at HEAD (one-piece shade kernels everywhere) + ncu of k_shade<Lambert, path> of bounce 0, frame 2, batch 4 + default bench
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
RUN="python tools/one_frame.py libpbrs_gpu.so c4 0.25 2"
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 327 -c 1 -f -o gpurun_out/r2_p31_shade $RUN > gpurun_out/r2_p31_ncu.log 2>&1; tail -2 gpurun_out/r2_p31_ncu.log
python bench.py --no-cpu > gpurun_out/r2_bench_c4_d.json 2> gpurun_out/r2_bench_c4_d.err; tail -2 gpurun_out/r2_bench_c4_d.err; cat gpurun_out/r2_bench_c4_d.json
