mkdir -p gpurun_out
for cfg in "2048 2" "4096 2" "4096 1" "2048 3" "8192 1"; do
  set -- $cfg
  timeout 300 python bench.py --stage render --steps 5 --warmup 3 --no-cpu-baseline --chunk-rays $1 --lanes $2 > gpurun_out/sweep_$1_$2.log 2>&1
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/sweep_{sys.argv[1]}_{sys.argv[2]}.log').read().strip().splitlines()[-1])
    print('chunk', sys.argv[1], 'lanes', sys.argv[2], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'gemm1 ms', round(d['roofline']['avg_launch_ms'], 4))
except Exception as e:
    print('chunk', sys.argv[1], 'lanes', sys.argv[2], 'failed', e, open(f'gpurun_out/sweep_{sys.argv[1]}_{sys.argv[2]}.log').read()[-400:])
PY
done
