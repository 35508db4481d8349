set -x
timeout 900 python -m pytest tests/test_gpu_host.py -x -q -k "simple_examples" > gpurun_out/r2p_simple_examples.log 2>&1; echo "simple rc=$?"; tail -25 gpurun_out/r2p_simple_examples.log
