mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
SLCU_RESOLVE=1 timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash tools/gpu_ab.sh "c3_indoor2m c2_instancing10k c4_skinned100 c1_sphere100k" base 2>&1 | tee gpurun_out/ab_r02l.txt
SLCU_DEBUG=1 timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --config c3_indoor2m 2>&1 >/dev/null | grep -E "resolve" | tail -2 | tee -a gpurun_out/ab_r02l.txt
