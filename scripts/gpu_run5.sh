set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_gather_gpu.py -x -q -m gpu > gpurun_out/r2_t_render7.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_render7.log | head
CPN_GATHER_SEQ=1 timeout 600 python -m pytest tests/test_gather_gpu.py tests/test_render_gpu.py -x -q -m gpu -k "gather or matches_reference_golden" > gpurun_out/r2_t_seq.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_seq.log | head
for sk in 0 1 2 3 4 7; do echo "== GEMM1 drain attribution, skip=$sk"; CPN_TC_DBG_SKIP=$sk timeout 120 python scripts/gemm1_trace.py 524288 0 1 2>&1 | grep -A1 -E "drain_us|tile_period|mainloop" | grep -E "us\"|mean"; done 2>&1 | tee gpurun_out/r2_drain_attribution.log
