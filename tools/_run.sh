mkdir -p gpurun_out
NN=${1:-8}; TAG=${2:-r02y}; N=$NN
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/gpu_sortfirst_check.py > gpurun_out/${TAG}_sortfirst_check_x$N.txt 2>&1
grep -E "ok |BAD|SORT-FIRST|Error|error" gpurun_out/${TAG}_sortfirst_check_x$N.txt | tail -16
for spec in "1 c3_indoor2m" "0 c3_indoor2m" "1 c2_instancing10k" "0 c2_instancing10k" "1 c5_volume8k"; do
  set -- $spec; share=$1; c=$2
  steps=10; [ $c = c5_volume8k ] && steps=4
  SLCU_GEOM_SHARE=$share timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps $steps --warmup 3 --no-cpu-baseline --no-other-configs --config $c > gpurun_out/${TAG}_scale_${c}_x${N}_share$share.json 2> gpurun_out/${TAG}_scale_${c}_x${N}_share$share.err
  python - gpurun_out/${TAG}_scale_${c}_x${N}_share$share.json <<'PY' || grep -v Warn gpurun_out/${TAG}_scale_${c}_x${N}_share$share.err | grep -i "error" | head -5
import json, sys
d = json.loads(open(sys.argv[1]).read().splitlines()[-1])
print(sys.argv[1], 'N', d['n_gpus'], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), {a: round(b, 3) for a, b in d['roofline']['kernel_ms_per_frame'].items()}, flush=True)
PY
done 2>&1 | tee gpurun_out/${TAG}_share_ab_x$N.txt
