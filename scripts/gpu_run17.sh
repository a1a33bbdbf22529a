set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py -x -q -m gpu > gpurun_out/r2_t_render9.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_render9.log | head
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_mlp16.log 2>&1; tail -c 1400 gpurun_out/r2_bench_mlp16.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_mlp16.csv python bench.py --stage render --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp16.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_mlp16.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_getz2.csv python scripts/bench_getz.py --iters 1 > gpurun_out/ncu_getz2.log 2>&1; tail -1 gpurun_out/ncu_getz2.log
