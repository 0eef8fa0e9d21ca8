# round-1 (h): ncu of k_resolve / k_span on C3
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 3 -c 1 -o gpurun_out/resolve_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_span -s 6 -c 2 -o gpurun_out/span_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resolve -s 3 -c 1 -o gpurun_out/resolve_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c2_instancing10k > gpurun_out/ncu_full_c2.log 2>&1
SLCU_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --config c3_indoor2m 2>&1 | grep "slcu\]" | head -3
ls -la gpurun_out/*.ncu-rep
