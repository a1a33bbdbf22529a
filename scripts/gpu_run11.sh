set -x
mkdir -p gpurun_out
timeout 300 python scripts/gemm1_bench.py 524288 5 0 256 > gpurun_out/r2_gemm1_variants4.log 2>&1; cat gpurun_out/r2_gemm1_variants4.log
timeout 120 python scripts/gemm1_trace.py 524288 0 1 > gpurun_out/r2_gemm1_trace_splitring.json 2>&1; python -c "
import json
d=json.load(open('gpurun_out/r2_gemm1_trace_splitring.json'))
print({k:(round(v['mean'],2) if isinstance(v,dict) else v) for k,v in d.items()})"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t_all.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_all.log | head
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_splitring.log 2>&1; tail -c 1600 gpurun_out/r2_bench_splitring.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lanes 3 > gpurun_out/r2_bench_splitring_l3.log 2>&1; tail -c 1600 gpurun_out/r2_bench_splitring_l3.log | head -c 400
