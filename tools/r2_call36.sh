# round 2, GPU call 36: parity with three lanes; paths in flight 8 / 16 / 32 Mi on the full-size default bench
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for p in 8388608 16777216 33554432; do python bench.py --no-cpu --steps 2 --warmup 2 --paths-in-flight $p 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pif $p', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), d['film_crc32'])"; done | tee gpurun_out/r2_pif_full.log
