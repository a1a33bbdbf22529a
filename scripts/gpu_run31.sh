mkdir -p gpurun_out
show='import sys, json
d = json.loads(sys.stdin.read())
print({k: (round(v["mean"], 2) if isinstance(v, dict) else v) for k, v in d.items() if k in ("untraced_launch_ms", "mainloop_issue_us", "drain_us", "tile_period_us", "ring_wait_at_tile_start_us")})'
timeout 600 python -m pytest tests/test_render_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== KG compact"; timeout 120 python scripts/gemm1_trace.py 524288 10 1 1 2>&1 | python -c "$show"
echo "== KG full"; timeout 120 python scripts/gemm1_trace.py 524288 10 1 0 2>&1 | python -c "$show"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'serial', round(d['serial']['value']), round(d['serial']['e2e']['value']), 'gemm1', round(d['roofline']['avg_launch_ms'], 4))"
