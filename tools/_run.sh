mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -5
timeout 600 python tools/gpu_prims_bench.py --only points_4m_1080p,points_4m_1080p_alpha,wire_2k_large_1080p 2>&1 | grep -v Warn | tee gpurun_out/r03b_prims_bench.txt | tail -12
