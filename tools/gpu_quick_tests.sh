mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error|FAILED" gpurun_out/pytest_gpu.log | tail -8
