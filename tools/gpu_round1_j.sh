mkdir -p gpurun_out
for c in c3_indoor2m c2_instancing10k c1_sphere100k c4_skinned100; do SLCU_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --config $c 2>&1 | grep "slcu\]" | head -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 3 -c 1 -o gpurun_out/resolve_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3.log 2>&1
