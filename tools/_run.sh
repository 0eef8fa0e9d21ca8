mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -5
for rb in 1 4 8; do
for c in c3_indoor2m c2_instancing10k; do
  SLCU_READ_BANDS=$rb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --config $c > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "read_bands=$rb" "$c" <<'PY' || tail -3 gpurun_out/ab.err
import json, sys
d = json.loads(open('gpurun_out/ab.json').read().splitlines()[-1])
print(sys.argv[1], sys.argv[2], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), 'pipelined', round(d['e2e']['pipelined']['ms_per_step'], 3), flush=True)
PY
done; done 2>&1 | tee gpurun_out/ab_r03a.txt
