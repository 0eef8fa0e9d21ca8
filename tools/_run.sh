mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/b.json 2> gpurun_out/b.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/b.json').read().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
PY
