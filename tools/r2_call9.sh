# round 2, GPU call 9 (2 GPUs): parity, the one-call multi-device path, torchrun bench at N = 2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
PYTHONPATH=. timeout 600 python tools/multi_gpu_probe.py c4 0.5 2 2 2>&1 | tail -4
PYTHONPATH=. timeout 600 python tools/multi_gpu_probe.py c5 0.125 2 2 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_c4_n2.json 2> gpurun_out/r2_bench_c4_n2.err; tail -3 gpurun_out/r2_bench_c4_n2.err; cat gpurun_out/r2_bench_c4_n2.json
PBRS_BENCH_WORKLOAD=c5 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --frame-scale 0.25 > gpurun_out/r2_bench_c5q_n2.json 2> gpurun_out/r2_bench_c5q_n2.err; tail -3 gpurun_out/r2_bench_c5q_n2.err; cat gpurun_out/r2_bench_c5q_n2.json
