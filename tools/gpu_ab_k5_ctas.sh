cd $GRAFT_REPO_ROOT
for c in 5 6 8; do
  CMB_K5_CTAS=$c timeout 200 python bench.py --workload mica --no-cpu-baseline --steps 6 > gpurun_out/r2w_mica_ctas$c.log 2>&1
  echo "CTAS=$c $(tail -1 gpurun_out/r2w_mica_ctas$c.log | grep -o '"ms_per_step": [0-9.]*' | head -1) $(tail -1 gpurun_out/r2w_mica_ctas$c.log | grep -o '"kernel_ms_per_step.*')"
done
