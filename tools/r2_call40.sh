# round 2, GPU call 40: parity soak of the round-2 kernels against the oracle on randomised scenes (both generator families + the material zoo)
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
timeout 600 python tools/soak_gpu.py 20000 21200 2>&1 | tail -3 | tee gpurun_out/r2_soak.log
timeout 400 python tools/soak_gpu.py 3000 3500 zoo 2>&1 | tail -3 | tee -a gpurun_out/r2_soak.log
