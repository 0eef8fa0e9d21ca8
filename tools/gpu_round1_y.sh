mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -5
tests/cpp/_build/sl_draw_test_b200
