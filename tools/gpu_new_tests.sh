set -x
timeout 900 python -m pytest tests/test_gpu_stats.py tests/test_gpu_chain.py tests/test_gpu_cluster.py tests/test_gpu_inter.py -x -q > gpurun_out/r2q_sim_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2q_sim_tests.log
bash tools/ab.sh "CMB_X=0" "CMB_NULL_DEDUP=0"
