mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; grep '^{"metric"' gpurun_out/r2_bench_n$N.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('n', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'], 1), 'serial', round(d['serial']['value']), 'strong', round(d['strong']['value']), round(d['strong']['ms_per_step'], 2), d['strong']['bit_exact_vs_single_gpu_render'])" || tail -30 gpurun_out/r2_bench_n$N.log
