mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -5
run() {
  for c in $CONFIGS; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/b.json 2> gpurun_out/b.err
    python -c "
import json,sys
d=json.load(open('gpurun_out/b.json'))
print('$1', '$c', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_frame'].items() if k in ('raster','shade','vertex')})
" || tail -3 gpurun_out/b.err
  done
}
CONFIGS="c2_instancing10k c3_indoor2m c1_sphere100k c4_skinned100 c1_sphere1m_4k"
run skew
SLCU_TICKET_SKEW=0 run rowmajor
SLCU_TICKET_SKEW=100 run skew100
CONFIGS="c3_indoor2m c2_instancing10k c4_skinned100"
SLCU_SEG_WIDTH=256 run skew_seg256
SLCU_SEG_WIDTH=128 run skew_seg128
for c in c3_indoor2m; do SLCU_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --config $c 2>&1 | grep "slcu\] resolve" | head -1; done
