# round-1 (b): parity on the GPU + bench of C2/C3/C1 + ncu launch list + full captures of the two top kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' >> gpurun_out/box.txt
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --config c3_indoor2m > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --config c1_sphere100k > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c2.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raster -s 3 -c 1 -o gpurun_out/raster_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vertex -s 3 -c 1 -o gpurun_out/vertex_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raster -s 800 -c 1 -o gpurun_out/raster_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c2.json gpurun_out/bench_c3.json gpurun_out/bench_c1.json
