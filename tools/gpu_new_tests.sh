set -x
timeout 900 python -m pytest tests/test_gpu_stats.py tests/test_gpu_cluster.py -x -q -k "simulat or null" > gpurun_out/r2q_sim_tests.log 2>&1; echo "sim rc=$?"; tail -5 gpurun_out/r2q_sim_tests.log
bash tools/ab.sh "CMB_X=0"
