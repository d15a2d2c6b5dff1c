# round 2, GPU call 35: 2 / 3 / 4 batch lanes on the full-size default bench
cd $GRAFT_REPO_ROOT
for l in libpbrs_gpu.so libv_l3.so libv_l4.so; do PBRS_GPU_LIB=$PWD/pbrs_b200/lib/$l python bench.py --no-cpu --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$l', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), d['film_crc32'])"; done | tee gpurun_out/r2_lanes_full.log
