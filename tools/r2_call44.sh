# round 2, GPU call 44 (4 GPUs): the default bench under torchrun at N = 4 at HEAD (completes the builder-run 1 / 2 / 4 / 8 curve)
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_final_bench_c4_n4.json 2> gpurun_out/r2_final_bench_c4_n4.err; tail -2 gpurun_out/r2_final_bench_c4_n4.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final_bench_c4_n4.json')); print('N=4', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), d['film_crc32'], d['e2e']['film_crc32'])"
