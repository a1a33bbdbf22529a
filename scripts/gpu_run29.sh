mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/ab_$tag.log 2>&1; python - $tag <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/ab_{sys.argv[1]}.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'serial', round(d['serial']['value']), 'gemm1', round(d['roofline']['avg_launch_ms'], 4))
PY
}
EXTRA="" run base A=1
EXTRA="" run epi24 CPN_TC_EPI_WARPS=24
EXTRA="--lanes 3" run lanes3 A=1
EXTRA="--lanes 3" run lanes3_epi24 CPN_TC_EPI_WARPS=24
EXTRA="" run base2 A=1
