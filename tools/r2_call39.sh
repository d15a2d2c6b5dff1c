# round 2, GPU call 39 (8 GPUs): the default bench under torchrun at N = 8 at HEAD + the one-call num_gpus = 8 path
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_final_bench_c4_n8.json 2> gpurun_out/r2_final_bench_c4_n8.err; tail -2 gpurun_out/r2_final_bench_c4_n8.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench_c4_n8.json')); print('N=8', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), d['film_crc32'], d['e2e']['film_crc32'])"
PYTHONPATH=. timeout 300 python tools/multi_gpu_probe.py c4 1.0 8 2 2>&1 | tail -2 | tee gpurun_out/r2_final_probe_c4_n8.log
