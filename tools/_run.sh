mkdir -p gpurun_out
echo "== default"; timeout -k 10 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
bash tools/gpu_ab.sh "c3_indoor2m c2_instancing10k c1_sphere100k c4_skinned100 c5_volume4k" base 2>&1 | tee gpurun_out/ab_r02f.txt
