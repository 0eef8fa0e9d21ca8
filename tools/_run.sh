mkdir -p gpurun_out
echo "== resolve2 forced"; SLCU_RESOLVE=2 timeout -k 10 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== default (heuristic)"; timeout -k 10 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
for r in 1 2; do
for c in c3_indoor2m c2_instancing10k c1_sphere100k c4_skinned100; do
SLCU_RESOLVE=$r timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/ab.json 2> gpurun_out/ab.err
python - "resolve$r" "$c" <<'PY' || tail -3 gpurun_out/ab.err
import json, sys
d = json.loads(open('gpurun_out/ab.json').read().splitlines()[-1])
k = d['roofline']['kernel_ms_per_frame']
print(sys.argv[1], sys.argv[2], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), {a: round(b, 3) for a, b in k.items()}, flush=True)
PY
done; done 2>&1 | tee gpurun_out/ab_r02b.txt
SLCU_RESOLVE=2 SLCU_DEBUG=1 timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m 2>&1 >/dev/null | grep resolve | tail -1 | tee -a gpurun_out/ab_r02b.txt
