# round-1 (i): hand-overs carry the registers (no accumulator replay); segment width / occupancy sweep of k_resolve
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
run() {
  for c in $CONFIGS; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/b.json 2> gpurun_out/b.err
    python -c "
import json,sys
d=json.load(open('gpurun_out/b.json'))
print('$1', '$c', round(d['ms_per_step'],3),'ms', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_frame'].items() if k in ('raster','shade','span_count','span_fill')})
" || tail -3 gpurun_out/b.err
  done
}
CONFIGS="c2_instancing10k c3_indoor2m c1_sphere100k c4_skinned100 c5_volume4k"
run base
for sw in 64 256 512 1024 4096; do SLCU_SEG_WIDTH=$sw run seg$sw; done
CONFIGS="c2_instancing10k c3_indoor2m c4_skinned100"
SLCU_NVCC_EXTRA="-DSL_RASTER_MIN_BLOCKS=4" python softlight_b200/build.py --force > /dev/null 2>&1
run mb4
SLCU_SEG_WIDTH=256 run mb4seg256
SLCU_NVCC_EXTRA="-DSL_RASTER_MIN_BLOCKS=8" python softlight_b200/build.py --force > /dev/null 2>&1
run mb8
