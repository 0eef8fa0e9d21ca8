# One parametrised evidence run on the GPU box (replaces the per-session scripts of round 1):
#   tools/gpu_evidence.sh <tag> [tests] [bench] [ncu] [launches]
# writes gpurun_out/<tag>_*; summaries are copied into profiles/ by hand (tools/ncu_summary.py, tools/ncu_counters.py).
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,driver_version --format=csv,noheader > gpurun_out/${TAG}_box.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${TAG}_box.txt
for what in "$@"; do
case $what in
tests)
  (time timeout -k 10 1500 python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.txt;;
bench)
  timeout -k 10 1200 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -c 600 gpurun_out/${TAG}_bench_default.json
  timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref_default.json 2>&1;;
launches)
  for c in c3_indoor2m c2_instancing10k; do
    timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${c%%_*}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --config $c > /dev/null 2>&1
  done;;
ncu|ncu1|ncu2)
  # (a call's outputs come back only while they stay under 64 MiB: the captures are split over two calls)
  SPECS="c3_indoor2m:k_resolve2:resolve2_c3 c3_indoor2m:k_shade:shade_c3 c2_instancing10k:k_vertex:vertex_c2"
  [ $what = ncu2 ] && SPECS="c3_indoor2m:^k_span$:span_c3 c2_instancing10k:k_resolve2:resolve2_c2 c3_indoor2m:k_vertex:vertex_c3"
  for spec in $SPECS; do
    IFS=: read cfg kern out <<< "$spec"
    timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 6 -c 1 -f -o gpurun_out/${TAG}_$out python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --config $cfg > gpurun_out/${TAG}_ncu_$out.log 2>&1
    echo "$out rc=$?"
  done;;
esac
done
