set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py -x -q -m gpu > gpurun_out/r2_t_render11.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_render11.log | head
for c in 0 1; do
  echo "== persist2 GEMM1 compact=$c"
  timeout 120 python scripts/gemm1_trace.py 524288 0 1 $c 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items()})"
done
echo "== persist2 GEMM1 compact=1 W full"
CPN_TC_W_FULL=1 timeout 120 python scripts/gemm1_trace.py 524288 0 1 1 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items()})"
echo "== persist2 KG"
timeout 120 python scripts/gemm1_trace.py 524288 10 1 1 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items()})"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_p2.log 2>&1; tail -c 1500 gpurun_out/r2_bench_p2.log
