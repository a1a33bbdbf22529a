set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2_t_multi_gpu.log 2>&1; tail -5 gpurun_out/r2_t_multi_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.log 2>&1; tail -c 4500 gpurun_out/r2_bench_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_bench_n2_ref.log 2>&1; tail -c 600 gpurun_out/r2_bench_n2_ref.log
