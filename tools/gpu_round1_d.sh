# round-1 (d): C4/C5 device shaders -- parity on the GPU + first bench of C4 / C5
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' >> gpurun_out/box.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --config c4_skinned100 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config c5_volume4k > gpurun_out/bench_c5_4k.json 2> gpurun_out/bench_c5_4k.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config c5_volume8k > gpurun_out/bench_c5_8k.json 2> gpurun_out/bench_c5_8k.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config c3_indoor2m > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c4.json gpurun_out/bench_c5_4k.json gpurun_out/bench_c5_8k.json; tail -3 gpurun_out/*.err
