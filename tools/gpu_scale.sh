# N-GPU evidence: sort-first composite check against the goldens + the bench line of the default config and C2 / C5-8K
#   tools/gpu_scale.sh <tag> <N>
TAG=$1; N=$2
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/gpu_sortfirst_check.py > gpurun_out/${TAG}_sortfirst_check_x$N.txt 2>&1
grep -E "ok |BAD|SORT-FIRST|Error|error" gpurun_out/${TAG}_sortfirst_check_x$N.txt | tail -20
for c in c3_indoor2m c2_instancing10k c5_volume8k; do
  steps=10; [ $c = c5_volume8k ] && steps=4
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps $steps --warmup 3 --no-cpu-baseline --no-other-configs --config $c > gpurun_out/${TAG}_scale_${c}_x$N.json 2> gpurun_out/${TAG}_scale_${c}_x$N.err
  python - gpurun_out/${TAG}_scale_${c}_x$N.json <<'PY' || tail -5 gpurun_out/${TAG}_scale_${c}_x$N.err
import json, sys
d = json.loads(open(sys.argv[1]).read().splitlines()[-1])
print(sys.argv[1], 'N', d['n_gpus'], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), {a: round(b, 3) for a, b in d['roofline']['kernel_ms_per_frame'].items()}, flush=True)
PY
done
