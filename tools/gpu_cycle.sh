timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in c1 c3; do python bench.py --workload $w --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value'],1), 'Msamples/s', round(d['mrays_per_s'],1), 'Mrays/s frac', round(d['roofline']['frac'],3), {k:round(v,2) for k,v in d['stages_ms'].items()})"; done
for w in c4 c5; do sc=0.25; [ $w = c5 ] && sc=0.125; python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --frame-scale $sc > gpurun_out/bench_${w}s.json 2> gpurun_out/bench_${w}s.err; tail -2 gpurun_out/bench_${w}s.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${w}s.json')); print('${w}s', round(d['value'],1), 'Msamples/s', round(d['mrays_per_s'],1), 'Mrays/s frac', round(d['roofline']['frac'],3), {k:round(v,2) for k,v in d['stages_ms'].items()})"; done
