# round-1 (q): sort-first across N GPUs of one box (run with gpurun --gpus N): bench per config + bit-exact check of the composite
mkdir -p gpurun_out
N=${N:-2}
for c in c2_instancing10k c3_indoor2m c5_volume8k; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --config $c > gpurun_out/bench_${c}_x$N.json 2> gpurun_out/bench_${c}_x$N.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_${c}_x$N.json'))
print('$c x$N', round(d['ms_per_step'],3),'ms', round(d['value'],1), d['unit'], 'e2e ms', round(d['e2e']['ms_per_step'],3))
" || tail -5 gpurun_out/bench_${c}_x$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/gpu_sortfirst_check.py > gpurun_out/sortfirst_check_x$N.log 2>&1; tail -5 gpurun_out/sortfirst_check_x$N.log
