mkdir -p gpurun_out
show='import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v["mean"], 2) if isinstance(v, dict) else v) for k, v in d.items() if k in ("untraced_launch_ms", "mainloop_issue_us", "drain_us", "tile_period_us", "ring_wait_at_tile_start_us")})'
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
for l in 0 10 8; do
  echo "== MMA only (no copies), layer $l"
  CPN_TC_DBG_SKIP=48 timeout 120 python scripts/gemm1_trace.py 524288 $l 1 2>&1 | python -c "$show"
done
echo "== nothing (no copies, no MMAs), layer 10"
CPN_TC_DBG_SKIP=56 timeout 120 python scripts/gemm1_trace.py 524288 10 1 2>&1 | python -c "$show"
echo "== copies only, layer 10"
CPN_TC_DBG_SKIP=8 timeout 120 python scripts/gemm1_trace.py 524288 10 1 2>&1 | python -c "$show"
