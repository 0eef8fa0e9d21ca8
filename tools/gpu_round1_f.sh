# round-1 (f): NaN-exact min/max; C5 parity; ncu full captures of the two-phase k_raster (C3, C2)
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -k "c5 or c4 or c3_small or rand_blend") > gpurun_out/pytest_gpu_c5.log 2>&1
tail -4 gpurun_out/pytest_gpu_c5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raster -s 3 -c 1 -o gpurun_out/raster_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raster -s 3 -c 1 -o gpurun_out/raster_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c2_instancing10k > gpurun_out/ncu_full_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
