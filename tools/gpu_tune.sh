# tuning sweeps: SLCU_NVCC_EXTRA builds on the box, bench of the listed configs
mkdir -p gpurun_out
run() {
  for c in $CONFIGS; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/b.json 2> gpurun_out/b.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/b.json').read().splitlines()[-1])
print('$1', '$c', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_frame'].items() if k in ('raster','shade','vertex')})
" || tail -3 gpurun_out/b.err
  done
}
CONFIGS="c2_instancing10k c3_indoor2m c4_skinned100"
for mb in 5 6; do
SLCU_NVCC_EXTRA="-DSL_RASTER_MIN_BLOCKS=$mb" python softlight_b200/build.py --force > /dev/null 2>&1
run rmb$mb
done
