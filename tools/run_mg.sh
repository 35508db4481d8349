cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -2 gpurun_out/bench_n2.log | cut -c1-1800
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-600
