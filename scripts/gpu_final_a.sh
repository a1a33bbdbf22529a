mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 700 gpurun_out/r2_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_render_default.csv python bench.py --stage render --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final_list.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_render_default.csv
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|gather_image|readout_image' -s 483 -c 6 -o gpurun_out/r2_chunk_full_final python bench.py --stage render --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_chunk_full_final.log 2>&1; tail -2 gpurun_out/ncu_chunk_full_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_getz_final.csv python scripts/bench_getz.py --iters 1 > gpurun_out/ncu_getz_final.log 2>&1; tail -1 gpurun_out/ncu_getz_final.log
timeout 300 python scripts/bench_getz.py --iters 20 2>&1 | tail -1 > gpurun_out/r2_getz_parts.json; cat gpurun_out/r2_getz_parts.json
