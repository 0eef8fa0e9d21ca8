# compute-sanitizer (memcheck, then racecheck on shared memory) over a subset of the parity cases
# (not the shared-vertex-stage test: the sanitizer serialises kernels, and that test's contexts wait for each other on the device)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_blit.py -m gpu -q -x -k "(c1_small or c2_small or c3_small or c4_small or c5_small or rand_clip_heavy or overflow_blend or overflow_half_ge or mrt_mixed or rand_blend1 or tiny_1x1 or ragged or lines_ or rand_index_u8 or present or blit_to_host) and not shared_vertex" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | tail -8
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_cuda_matches_oracle and (c2_small or c3_small or rand_clip_heavy or overflow_blend)" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | tail -8
