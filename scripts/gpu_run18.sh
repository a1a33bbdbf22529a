set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ufc_native_gpu.py tests/test_ufc_gpu.py tests/test_pair_gpu.py -x -q -m gpu > gpurun_out/r2_t_ufc10.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_ufc10.log | head
timeout 300 python scripts/bench_getz.py --iters 20 > gpurun_out/getz3.log 2>&1; tail -3 gpurun_out/getz3.log
CPN_CONV4D_DIRECT=1 timeout 300 python scripts/bench_getz.py --iters 20 > gpurun_out/getz3_direct.log 2>&1; tail -3 gpurun_out/getz3_direct.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_getz3.csv python scripts/bench_getz.py --iters 1 > gpurun_out/ncu_getz3.log 2>&1; tail -1 gpurun_out/ncu_getz3.log
python scripts/summarize_launches.py gpurun_out/r2_launches_getz3.csv | head -40
