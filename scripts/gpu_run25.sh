mkdir -p gpurun_out
show='import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v["mean"], 2) if isinstance(v, dict) else v) for k, v in d.items() if k in ("untraced_launch_ms", "mainloop_issue_us", "drain_us", "tile_period_us", "ring_wait_at_tile_start_us")})'
timeout 600 python -m pytest tests/test_render_gpu.py -x -q -m gpu 2>&1 | tail -3
for l in 0 10 8; do
  echo "== MMA only (no copies), layer $l"
  CPN_TC_DBG_SKIP=48 timeout 120 python scripts/gemm1_trace.py 524288 $l 1 2>&1 | python -c "$show"
  echo "== full, layer $l"
  timeout 120 python scripts/gemm1_trace.py 524288 $l 1 2>&1 | python -c "$show"
done
echo "== KG compact (default)"
timeout 120 python scripts/gemm1_trace.py 524288 10 1 1 2>&1 | python -c "$show"
echo "== KG compact WS"
CPN_TC_WS=1 timeout 120 python scripts/gemm1_trace.py 524288 10 1 1 2>&1 | python -c "$show"
echo "== GEMM1 prefetch 4"
CPN_TC_PREFETCH=4 timeout 120 python scripts/gemm1_trace.py 524288 0 1 2>&1 | python -c "$show"
echo "== GEMM1 compact A"
timeout 120 python scripts/gemm1_trace.py 524288 0 1 1 2>&1 | python -c "$show"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -c 1500
