# round-1 (final, refresh): bench of every config on the committed state (after the scan / harness changes)
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for c in c1_sphere100k c3_indoor2m c4_skinned100 c5_volume4k; do
  timeout 900 python bench.py --steps 10 --warmup 3 --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config c5_volume8k > gpurun_out/bench_c5_volume8k.json 2> gpurun_out/bench_c5_volume8k.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_default.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
for f in gpurun_out/bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().splitlines()[-1])
print('$f', d.get('impl','cuda'), round(d['ms_per_step'],3),'ms', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'))
"; done
