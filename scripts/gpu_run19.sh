set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_gather_gpu.py -x -q -m gpu > gpurun_out/r2_t_render10.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_render10.log | head
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a3.log 2>&1; tail -c 1400 gpurun_out/r2_bench_a3.log
CPN_TC_W_FULL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a3_wfull.log 2>&1; tail -c 700 gpurun_out/r2_bench_a3_wfull.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_a3.csv python bench.py --stage render --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a3.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_a3.csv
