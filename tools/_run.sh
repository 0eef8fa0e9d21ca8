mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -5
bash tools/gpu_ab.sh "c2_instancing10k c3_indoor2m c4_skinned100" base 2>&1 | tee gpurun_out/ab_r02w.txt
