mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 1500 gpurun_out/r2_bench_default.json
SEL='not full_resolution and not full_image and not config4 and not 512 and not two_gpu and not graph_replay and not batches and not 4096 and not pose_feature_operators and not render_pairs'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -x -q -m gpu -k "$SEL" -p no:cacheprovider > gpurun_out/r2_sanitizer_memcheck_final.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2_sanitizer_memcheck_final.log | tail -5
