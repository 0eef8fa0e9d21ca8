mkdir -p gpurun_out
run() {
  for c in $CONFIGS; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/b.json 2> gpurun_out/b.err
    python -c "
import json,sys
d=json.load(open('gpurun_out/b.json'))
print('$1', '$c', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_frame'].items() if k in ('vertex',)})
" || tail -3 gpurun_out/b.err
  done
}
CONFIGS="c2_instancing10k c3_indoor2m c4_skinned100 c1_sphere100k"
SLCU_VERTEX_CACHE=1 run cache
