# compute-sanitizer memcheck over the point and wireframe parity cases (sl_points.cuh, sl_wire.cuh)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_cuda_matches_oracle and (points_ or wire_)" > gpurun_out/sanitize_prims.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_prims.log | tail -8
