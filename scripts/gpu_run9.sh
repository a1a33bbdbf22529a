set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ufc_native_gpu.py tests/test_ufc_gpu.py tests/test_pair_gpu.py -x -q -m gpu > gpurun_out/r2_t_ufc2.log 2>&1; grep -E "passed|failed|FAILED|Error|rel err|oracle:" gpurun_out/r2_t_ufc2.log | head -30
for t in 2048 1000000000 64; do echo "== tc_min_rows $t"; CPN_UFC_TC_MIN_ROWS=$t timeout 120 python scripts/bench_getz.py --iters 10 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_render_gpu.py -x -q -m gpu -k "oblique_pose" > gpurun_out/r2_t_obl.log 2>&1; grep -E "passed|failed|FAILED|Error|full-image" gpurun_out/r2_t_obl.log | head -20 | cut -c1-300
