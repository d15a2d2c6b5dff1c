# round 2, GPU call 37 (2 GPUs): multi-GPU tests + torchrun bench at N = 2 at HEAD (three lanes, one-piece shade kernels)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_final_bench_c4_n2.json 2> gpurun_out/r2_final_bench_c4_n2.err; tail -2 gpurun_out/r2_final_bench_c4_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench_c4_n2.json')); print('N=2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['film_crc32'], d['e2e']['film_crc32'])"
