# round 2, GPU call 26: load balance of an 8-way tile split measured on one GPU, three tile deals (t % N, (tx + 3 ty) % N, (tx + 5 ty) % N)
cd $GRAFT_REPO_ROOT
export PYTHONPATH=. PBRS_GPU_LIB=$PWD/pbrs_b200/lib/libv_deal.so
for d in 0 1 2; do echo deal $d; PBRS_TILE_DEAL=$d timeout 600 python tools/balance_probe.py c4 1.0 8 2 2>&1 | tail -2; done | tee gpurun_out/r2_balance.log
