mkdir -p gpurun_out
show='import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v["mean"], 2) if isinstance(v, dict) else v) for k, v in d.items() if k in ("untraced_launch_ms", "mainloop_issue_us", "drain_us", "tile_period_us", "ring_wait_at_tile_start_us")})'
for pf in 0 2 4 8 16; do
  echo "== GEMM1 prefetch=$pf"
  CPN_TC_PREFETCH=$pf timeout 120 python scripts/gemm1_trace.py 524288 0 1 2>&1 | python -c "$show"
  echo "== GEMM1 prefetch=$pf all CTAs"
  CPN_TC_PREFETCH_ALL=1 CPN_TC_PREFETCH=$pf timeout 120 python scripts/gemm1_trace.py 524288 0 1 2>&1 | python -c "$show"
done
for pf in 0 4 8 16; do
  echo "== KG prefetch=$pf"
  CPN_TC_PREFETCH=$pf timeout 120 python scripts/gemm1_trace.py 524288 10 1 1 2>&1 | python -c "$show"
done
echo "== copies only, prefetch 8"
CPN_TC_DBG_SKIP=8 CPN_TC_PREFETCH=8 timeout 120 python scripts/gemm1_trace.py 524288 0 1 2>&1 | python -c "$show"
