# ncu evidence for the point and wireframe kernels: launch list of tools/gpu_prims_bench.py + full captures of the top kernels
mkdir -p gpurun_out
export SLCU_PRIMS_FRAMES=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_prims.csv python tools/gpu_prims_bench.py --no-ref > gpurun_out/ncu_prims.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_point_shade -c 1 -o gpurun_out/point_shade python tools/gpu_prims_bench.py --no-ref --only points_4m_1080p > gpurun_out/ncu_ps.log 2>&1
echo "point_shade rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wire_shade -c 1 -o gpurun_out/wire_shade python tools/gpu_prims_bench.py --no-ref --only wire_200k_1080p > gpurun_out/ncu_ws.log 2>&1
echo "wire_shade rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wire_depth -c 1 -o gpurun_out/wire_depth python tools/gpu_prims_bench.py --no-ref --only wire_200k_1080p > gpurun_out/ncu_wd.log 2>&1
echo "wire_depth rc=$?"
ls -la gpurun_out/*.ncu-rep
