# round-1 (g): fragment stage split into k_resolve + k_shade<FS>: parity + bench + launch list
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
for c in c2_instancing10k c3_indoor2m c1_sphere100k c4_skinned100 c5_volume4k c5_volume8k; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$c.json'))
print('$c', round(d['ms_per_step'],3),'ms', round(d['gfrags_per_s'],3),'Gfrag/s', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_frame'].items()})
" || tail -3 gpurun_out/bench_$c.err
done
