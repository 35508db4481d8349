set -x
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2q_bench_n8.log 2>&1; echo "n8 rc=$?"; tail -1 gpurun_out/r2q_bench_n8.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2q_bench_n4.log 2>&1; echo "n4 rc=$?"; tail -1 gpurun_out/r2q_bench_n4.log | cut -c1-300
