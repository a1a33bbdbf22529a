mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pair_gpu.py tests/test_ufc_native_gpu.py -x -q -m gpu 2>&1 | grep -E "rel_pose|passed|failed|Error" | tail -8
timeout 300 python scripts/bench_getz.py --iters 20 2>&1 | tail -1
CPN_UFC_SIMT_LINEAR=1 timeout 300 python scripts/bench_getz.py --iters 20 2>&1 | tail -1
