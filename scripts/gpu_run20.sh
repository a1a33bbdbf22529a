set -x
mkdir -p gpurun_out
for sk in 0 8 16 32 24 40 48 56; do
  echo "== GEMM1 main-loop attribution, skip=$sk"
  CPN_TC_DBG_SKIP=$sk timeout 120 python scripts/gemm1_trace.py 524288 0 1 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items() if k in ('untraced_launch_ms', 'mainloop_issue_us', 'drain_us', 'tile_period_us', 'ring_wait_at_tile_start_us')})"
done
for sk in 0 8; do
  echo "== compact A, skip=$sk"
  CPN_TC_DBG_SKIP=$sk timeout 120 python scripts/gemm1_trace.py 524288 0 1 1 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items() if k in ('untraced_launch_ms', 'mainloop_issue_us', 'drain_us', 'tile_period_us', 'ring_wait_at_tile_start_us')})"
  CPN_TC_W_FULL=1 CPN_TC_DBG_SKIP=$sk timeout 120 python scripts/gemm1_trace.py 524288 0 1 1 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('W full', {k: (round(v['mean'], 2) if isinstance(v, dict) else v) for k, v in d.items() if k in ('untraced_launch_ms', 'mainloop_issue_us', 'drain_us', 'tile_period_us', 'ring_wait_at_tile_start_us')})"
done
