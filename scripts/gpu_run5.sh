set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_render_gpu.py -x -q -m gpu > gpurun_out/r2_t_render5.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/r2_t_render5.log | head
for cfg in "16 2" "16 1" "16 3" "24 2" "24 1"; do set -- $cfg; echo "== epi_warps=$1 lanes=$2"; CPN_TC_EPI_WARPS=$1 timeout 300 python bench.py --stage render --steps 5 --warmup 3 --no-cpu-baseline --lanes $2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'gemm1 ms', d['roofline']['avg_launch_ms'], 'clk', d['clocks']['sm_mhz'])
"; done 2>&1 | tee gpurun_out/r2_lanes_overlap.log
timeout 120 python scripts/gemm1_trace.py 524288 10 1 > gpurun_out/r2_kg_trace_cb16.json 2>&1; grep -A1 -E "untraced|drain_us|tile_period" gpurun_out/r2_kg_trace_cb16.json | head -12
