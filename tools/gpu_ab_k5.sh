cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_mica.py -x -q -s > $O/${1}_mica_tests.log 2>&1; echo "mica rc=$?"; tail -4 $O/${1}_mica_tests.log
for c in 5 6 8; do
  CMB_K5_CTAS=$c timeout 300 python bench.py --workload mica --no-cpu-baseline --steps 5 > $O/${1}_mica_ctas$c.log 2>&1
  echo "CTAS=$c $(tail -1 $O/${1}_mica_ctas$c.log | grep -o '"kernel_ms_per_step.*')"
done
