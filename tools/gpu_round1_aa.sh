mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x -k "c5 or volume") > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -5
run() {
  for c in $CONFIGS; do
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/b.json 2> gpurun_out/b.err
    python -c "
import json,sys
d=json.load(open('gpurun_out/b.json'))
print('$1', '$c', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_frame'].items() if k in ('shade',)})
" || tail -3 gpurun_out/b.err
  done
}
CONFIGS="c5_volume4k c3_indoor2m c2_instancing10k c4_skinned100"
run base
for mb in 3 6 8; do
SLCU_NVCC_EXTRA="-DSL_SHADE_MIN_BLOCKS=$mb" python softlight_b200/build.py --force > /dev/null 2>&1
run smb$mb
done
