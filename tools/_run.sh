mkdir -p gpurun_out
for fr in 0 3; do
for c in c3_indoor2m; do
  SLCU_FINE_ROWS=$fr timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --config $c > gpurun_out/s.json 2> gpurun_out/s.err
  python - "fine_rows=$fr" <<'PY' || tail -5 gpurun_out/s.err
import json, sys
d = json.loads(open('gpurun_out/s.json').read().splitlines()[-1])
print(sys.argv[1], 'N', d['n_gpus'], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), {a: round(b, 3) for a, b in d['roofline']['kernel_ms_per_frame'].items()}, flush=True)
PY
done; done 2>&1 | tee gpurun_out/ab_r02o.txt
bash tools/gpu_scale.sh r02o 4
