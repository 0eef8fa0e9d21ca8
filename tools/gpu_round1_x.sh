# quick regression: full GPU parity + device-time bench of the main configs
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -5
run() {
  for c in $CONFIGS; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/b.json 2> gpurun_out/b.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/b.json').read().splitlines()[-1])
print('$1', '$c', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_frame'].items()})
" || tail -3 gpurun_out/b.err
  done
}
CONFIGS="c2_instancing10k c3_indoor2m c1_sphere100k c4_skinned100"
run head

