# mica's GPU tests, then the interval filter of k5_permutations on / off: same tables, kernel time of each
cd $GRAFT_REPO_ROOT
O=gpurun_out
T=${1:-r2w}
timeout 600 python -m pytest tests/test_gpu_mica.py -x -q -s > $O/${T}_mica_tests.log 2>&1; echo "mica rc=$?"; tail -4 $O/${T}_mica_tests.log
CMB_K5_FILTER=1 python tools/k5_filter_check.py /tmp/f1.npz > $O/${T}_filter_check.log 2>&1
CMB_K5_FILTER=0 python tools/k5_filter_check.py /tmp/f0.npz >> $O/${T}_filter_check.log 2>&1
python -c "
import numpy as np
a, b = np.load('/tmp/f1.npz'), np.load('/tmp/f0.npz')
for k in a.files: print(k, 'identical' if np.array_equal(a[k], b[k]) else 'DIFFERENT: %d rows' % int((a[k] != b[k]).sum()))
" >> $O/${T}_filter_check.log 2>&1
cat $O/${T}_filter_check.log
for f in 1 0; do
  CMB_K5_FILTER=$f timeout 300 python bench.py --workload mica --no-cpu-baseline --steps 5 > $O/${T}_mica_filter$f.log 2>&1
  echo "FILTER=$f $(tail -1 $O/${T}_mica_filter$f.log | grep -o '"kernel_ms_per_step.*')"
done
