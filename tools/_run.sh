mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -5
for vc in 1 0; do
for c in c3_indoor2m c4_skinned100 c2_instancing10k c1_sphere100k; do
  SLCU_VERTEX_CACHE=$vc timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --config $c > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "vcache=$vc" "$c" <<'PY' || tail -3 gpurun_out/ab.err
import json, sys
d = json.loads(open('gpurun_out/ab.json').read().splitlines()[-1])
k = d['roofline']['kernel_ms_per_frame']
print(sys.argv[1], sys.argv[2], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), {a: round(b, 3) for a, b in k.items()}, flush=True)
PY
done; done 2>&1 | tee gpurun_out/ab_r03d.txt
