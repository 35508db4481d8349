set -x
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_map.py tests/test_gpu_stats.py -x -q -k "simple_examples or ancestral or label" > gpurun_out/r2p_simple_examples.log 2>&1; echo "simple rc=$?"; tail -25 gpurun_out/r2p_simple_examples.log
