timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for hf in 0 1; do SLCU_HEAVY_FIRST=$hf bash tools/gpu_ab.sh "c3_indoor2m c2_instancing10k c4_skinned100" base 2>&1 | sed "s/^/heavy_first=$hf /" | tee -a gpurun_out/ab_r02h.txt; done
SLCU_DEBUG=1 timeout -k 10 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --config c3_indoor2m 2>&1 >/dev/null | grep "resolve:" | tail -1 | tee -a gpurun_out/ab_r02h.txt
bash tools/gpu_scale.sh r02h 2
