for pif in 1048576 4194304 16777216 33554432; do
for w in c3 c4; do sc=1.0; [ $w = c4 ] && sc=0.25; python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --frame-scale $sc --paths-in-flight $pif > gpurun_out/t.json 2> gpurun_out/t.err; tail -2 gpurun_out/t.err; python -c "
import json,sys; d=json.load(open('gpurun_out/t.json')); print('pif $pif $w', round(d['value'],1), 'Msamples/s', round(d['mrays_per_s'],1), 'Mrays/s frac', round(d['roofline']['frac'],3), {k:round(v,2) for k,v in d['stages_ms'].items()})"; done; done
