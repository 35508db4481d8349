cd $GRAFT_REPO_ROOT
timeout 55 python -m pytest tests/test_gpu_mica.py tests/test_gpu_host.py -x -q -k "mica or permutation or site_statistics or parametric" > gpurun_out/r2w_final_mica_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2w_final_mica_tests.log
