# round-1 (c): batched draws -- parity + bench of C2/C3/C1 + launch list and full capture for C3
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config c1_sphere100k > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raster -s 3 -c 1 -o gpurun_out/raster_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_span -s 6 -c 2 -o gpurun_out/span_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/ncu_full_c3s.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c2.json gpurun_out/bench_c3.json gpurun_out/bench_c1.json
