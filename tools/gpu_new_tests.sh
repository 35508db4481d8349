set -x
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_mica.py -x -q -k "mica" > gpurun_out/r2q_mica_tests.log 2>&1; echo "mica rc=$?"; tail -30 gpurun_out/r2q_mica_tests.log
