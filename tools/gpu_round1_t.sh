mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 3 -c 1 -o gpurun_out/resolve_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3.log 2>&1
