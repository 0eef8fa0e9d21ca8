mkdir -p gpurun_out
SLCU_RESOLVE=1 timeout -k 10 900 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_blit.py -m gpu -q -x -k "test_cuda_matches_oracle and (c1_small or c2_small or c3_small or c4_small or rand_clip_heavy or overflow_blend or overflow_half_ge or rand_blend1 or tiny_1x1 or ragged or lines_ or points_dense or wire_alpha or rand_index_u8) or shared_vertex and c2_small or present or blit_to_host" > gpurun_out/r03c_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r03c_sanitize_memcheck.log | tail -8
