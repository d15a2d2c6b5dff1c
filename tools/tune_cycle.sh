for cfg in "libpbrs_gpu_l1.so 16777216" "libpbrs_gpu.so 16777216" "libpbrs_gpu.so 8388608" "libpbrs_gpu.so 33554432"; do
set -- $cfg; lib=$1; pif=$2
export PBRS_GPU_LIB=$PWD/pbrs_b200/lib/$lib
for w in c3 c4 c5; do sc=1.0; [ $w = c4 ] && sc=0.5;  [ $w = c5 ] && sc=0.25; python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --frame-scale $sc --paths-in-flight $pif > gpurun_out/t.json 2> gpurun_out/t.err; tail -2 gpurun_out/t.err; python -c "
import json,sys; d=json.load(open('gpurun_out/t.json')); print('$lib $pif $w', round(d['value'],1), 'Msamples/s e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],1), 'sum stages', round(sum(v for k,v in d['stages_ms'].items() if k!='ms_total'),1))"; done; done
