set -x
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_stats.py -x -q -k "variants or label or laplace or count_methods" > gpurun_out/r2p_new_tests.log 2>&1; echo "new rc=$?"; tail -15 gpurun_out/r2p_new_tests.log
timeout 900 python -m pytest tests/test_gpu_host.py -x -q -k "label or laplace or mutual" > gpurun_out/r2p_new_host_tests.log 2>&1; echo "host rc=$?"; tail -15 gpurun_out/r2p_new_host_tests.log
