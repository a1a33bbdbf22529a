set -x
mkdir -p gpurun_out
timeout 120 python scripts/gemm1_trace.py 524288 0 2 > gpurun_out/r2_gemm1_trace_pair.json 2>&1; cat gpurun_out/r2_gemm1_trace_pair.json | python -c "
import sys, json
d = json.load(sys.stdin)
print({k: (v['mean'] if isinstance(v, dict) else v) for k, v in d.items()})"
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_config4.log 2>&1; tail -c 4500 gpurun_out/r2_bench_config4.log
timeout 120 python scripts/bench_getz.py --iters 10 > gpurun_out/r2_getz_parts.json 2>&1; tail -2 gpurun_out/r2_getz_parts.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_getz.csv python scripts/bench_getz.py --iters 1 > gpurun_out/ncu_getz.log 2>&1; tail -2 gpurun_out/ncu_getz.log
