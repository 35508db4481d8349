set -x
timeout 900 python -m pytest tests/test_gpu_mica.py -x -q > gpurun_out/r2q_mica_tests.log 2>&1; echo "mica rc=$?"; tail -25 gpurun_out/r2q_mica_tests.log
