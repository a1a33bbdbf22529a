mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pair_gpu.py tests/test_ufc_native_gpu.py tests/test_ufc_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_pipe.log 2>&1; tail -c 2600 gpurun_out/r2_bench_pipe.log
timeout 300 python scripts/bench_getz.py --iters 20 2>&1 | tail -1
