# round 2, GPU call 11: source-level ncu capture of the two traversal kernels at HEAD (bounce 1 of a mid-frame batch, C4 at 1/4 frame:
# same 16 Mi-path batches as the full frame) + of the shade kernels of that stage; summaries extracted on the box
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
python tools/one_frame.py libpbrs_gpu.so c4 0.25 2 > gpurun_out/r2_p11_plain.log 2>&1; cat gpurun_out/r2_p11_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 92 -c 2 -f -o gpurun_out/r2_p11_trace python tools/one_frame.py libpbrs_gpu.so c4 0.25 2 > gpurun_out/r2_p11_ncu1.log 2>&1; tail -2 gpurun_out/r2_p11_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:"k_surface|k_scatter" -s 18 -c 2 -f -o gpurun_out/r2_p11_shade python tools/one_frame.py libpbrs_gpu.so c4 0.25 2 > gpurun_out/r2_p11_ncu2.log 2>&1; tail -2 gpurun_out/r2_p11_ncu2.log
ls -la gpurun_out
