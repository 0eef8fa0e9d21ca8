# N-GPU scaling lines (one box): tools/gpu_scale.sh <tag> <N> [check]
#   the bench line of C3 (default: replicated vertex stage below 4 M primitives), C3 with the shared stage forced, C2, C5-8K;
#   with "check": the sort-first composite check against the goldens first (tools/gpu_sortfirst_check.py)
TAG=$1; N=$2
mkdir -p gpurun_out
if [ "$3" = check ]; then
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/gpu_sortfirst_check.py > gpurun_out/${TAG}_sortfirst_check_x$N.txt 2>&1
  grep -E "ok |BAD|SORT-FIRST|Error|error" gpurun_out/${TAG}_sortfirst_check_x$N.txt | tail -16
fi
for spec in "default c3_indoor2m" "shared c3_indoor2m" "default c2_instancing10k" "default c5_volume8k"; do
  set -- $spec; mode=$1; c=$2
  steps=10; [ $c = c5_volume8k ] && steps=4
  if [ $mode = shared ]; then export SLCU_GEOM_SHARE_MIN_PRIMS=0; else unset SLCU_GEOM_SHARE_MIN_PRIMS; fi
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps $steps --warmup 3 --no-cpu-baseline --no-other-configs --config $c > gpurun_out/${TAG}_scale_${c}_x${N}_$mode.json 2> gpurun_out/${TAG}_scale_${c}_x${N}_$mode.err
  python - gpurun_out/${TAG}_scale_${c}_x${N}_$mode.json $mode <<'PY' || grep -v Warn gpurun_out/${TAG}_scale_${c}_x${N}_$mode.err | grep -i "error" | head -5
import json, sys
d = json.loads(open(sys.argv[1]).read().splitlines()[-1])
print(sys.argv[2], d['config']['workload'][:16], 'N', d['n_gpus'], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), {a: round(b, 3) for a, b in d['roofline']['kernel_ms_per_frame'].items()}, flush=True)
PY
done 2>&1 | tee gpurun_out/${TAG}_scale_x$N.txt
