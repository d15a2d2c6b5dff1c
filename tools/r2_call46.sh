# round 2, GPU call 46: parity + default bench on the final build (multi-lobe routines inlined)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --no-cpu > gpurun_out/r2_final3_bench_c4.json 2> gpurun_out/r2_final3_bench_c4.err; tail -2 gpurun_out/r2_final3_bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/r2_final3_bench_c4.json')); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['film_crc32'], round(d['roofline']['frac'],3), {k[3:]:round(v,1) for k,v in d['stages_ms'].items()})"
