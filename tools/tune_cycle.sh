# A/B harness: usage  LIBS="libpbrs_gpu.so libX.so" WORKLOADS="c3 c4" bash tools/tune_cycle.sh
# (alternative builds: make -C pbrs_b200/csrc OUT=../lib/libX.so EXTRA=-DPBRS_...=v)
LIBS=${LIBS:-"libpbrs_gpu.so"}
WORKLOADS=${WORKLOADS:-"c3 c4 c5"}
for lib in $LIBS; do
export PBRS_GPU_LIB=$PWD/pbrs_b200/lib/$lib
for w in $WORKLOADS; do sc=1.0; [ $w = c4 ] && sc=0.25;  [ $w = c5 ] && sc=0.125; python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --frame-scale $sc > gpurun_out/t.json 2> gpurun_out/t.err; tail -2 gpurun_out/t.err; python -c "
import json,sys; d=json.load(open('gpurun_out/t.json')); print('$lib $w', round(d['value'],1), 'Msamples/s frac', round(d['roofline']['frac'],3), {k[3:]:round(v,1) for k,v in d['stages_ms'].items()})"; done; done
