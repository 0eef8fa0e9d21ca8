mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
SLCU_DEBUG=1 timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --config c3_indoor2m 2>&1 >/dev/null | grep -E "resolve" | tail -2
