set -x
mkdir -p gpurun_out
CPN_TC_WS=1 timeout 300 python -m pytest tests/test_render_gpu.py -x -q -m gpu -k "key_and_round2 or cb16_and_rowdot or folded_value" > gpurun_out/r2_t_ws1.log 2>&1; tail -5 gpurun_out/r2_t_ws1.log
CPN_TC_WS=1 timeout 600 python -m pytest tests/test_render_gpu.py -x -q -m gpu > gpurun_out/r2_t_ws2.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_ws2.log | head
for ws in 0 1; do
  echo "== KG ws=$ws"
  CPN_TC_WS=$ws timeout 120 python scripts/gemm1_trace.py 524288 10 1 1 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items()})"
  CPN_TC_WS=$ws timeout 120 python scripts/gemm1_trace.py 524288 10 1 0 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('full A', {k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items()})"
done
