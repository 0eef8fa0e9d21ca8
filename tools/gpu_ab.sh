# A/B of product-library variants built with softlight_b200/build.py --out=build/<name>.so:
#   tools/gpu_ab.sh "<configs>" <variant> [<variant> ...]      ("base" = softlight_b200/libslcu.so)
# prints ms/frame and the per-kernel split of every (variant, config); nothing here is a bench value.
mkdir -p gpurun_out
CONFIGS="$1"; shift
for v in "$@"; do
  lib=build/$v.so; [ "$v" = base ] && lib=softlight_b200/libslcu.so
  for c in $CONFIGS; do
    SLCU_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --config $c > gpurun_out/ab.json 2> gpurun_out/ab.err
    python - "$v" "$c" <<'PY' || tail -3 gpurun_out/ab.err
import json, sys
d = json.loads(open('gpurun_out/ab.json').read().splitlines()[-1])
k = d['roofline']['kernel_ms_per_frame']
print(sys.argv[1], sys.argv[2], round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['ms_per_step'], 3), {a: round(b, 3) for a, b in k.items()}, flush=True)
PY
  done
done
