# round 2, GPU call 24 (8 GPUs): the default bench under torchrun at N = 8 (C4, tile split, own tiles -> one shared page-locked film),
# full-size C5 (sample split, NCCL reduce) at N = 8, and the one-call num_gpus = 8 path of pbrs_render
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_c4_n8.json 2> gpurun_out/r2_bench_c4_n8.err; tail -3 gpurun_out/r2_bench_c4_n8.err; cat gpurun_out/r2_bench_c4_n8.json
PBRS_BENCH_WORKLOAD=c5 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 1 --warmup 1 > gpurun_out/r2_bench_c5_n8.json 2> gpurun_out/r2_bench_c5_n8.err; tail -3 gpurun_out/r2_bench_c5_n8.err; cat gpurun_out/r2_bench_c5_n8.json
PYTHONPATH=. timeout 600 python tools/multi_gpu_probe.py c4 1.0 8 2 2>&1 | tail -4 | tee gpurun_out/r2_probe_c4_n8.log
