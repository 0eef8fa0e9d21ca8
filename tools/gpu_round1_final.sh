# round-1 (final, r01g): evidence run for the resolve/shade pipeline: parity log, bench of every config (with the CPU reference beside it),
# reference arm, ncu launch lists and full captures of the top kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' >> gpurun_out/box.txt
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for c in c1_sphere100k c3_indoor2m c4_skinned100 c5_volume4k; do
  timeout 900 python bench.py --steps 10 --warmup 3 --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config c5_volume8k > gpurun_out/bench_c5_volume8k.json 2> gpurun_out/bench_c5_volume8k.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_default.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_l3.log 2>&1
cap() { # name kernel-regex config skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -o gpurun_out/$1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config $3 > gpurun_out/ncu_$1.log 2>&1
}
cap resolve_c2 k_resolve c2_instancing10k 3
cap vertex_c2 k_vertex c2_instancing10k 3
cap resolve_c3 k_resolve c3_indoor2m 3
cap shade_c3 k_shade c3_indoor2m 3
cap shade_c5 k_shade c5_volume4k 3
for f in gpurun_out/bench_*.json; do python -c "
import json,sys
d=json.load(open('$f'))
print('$f', d.get('impl','cuda'), round(d['ms_per_step'],3),'ms', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'))
"; done
cap span_c3 k_span c3_indoor2m 6
for c in c3_indoor2m c2_instancing10k; do SLCU_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --config $c 2>&1 | grep "slcu\] resolve\|slcu\] batch" | head -2; done > gpurun_out/resolve_diag.txt
cat gpurun_out/resolve_diag.txt
