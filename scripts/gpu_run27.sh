mkdir -p gpurun_out
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_pipe2.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_pipe2.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','serial')}, d['e2e'])
PY
CPN_PIPE_PRIORITY=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_pipe3.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_pipe3.log').read().strip().splitlines()[-1])
print('normal priority', {k:d[k] for k in ('value','ms_per_step','serial')}, d['e2e'])
PY
