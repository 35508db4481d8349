set -x
cd $GRAFT_REPO_ROOT
T=${1:-r2r}
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_mica.py -x -q -s > $O/${T}_mica_tests.log 2>&1; echo "mica rc=$?"; tail -8 $O/${T}_mica_tests.log
timeout 900 python -m pytest tests/test_gpu_host.py -x -q -k "mica" > $O/${T}_mica_cli_tests.log 2>&1; echo "mica cli rc=$?"; tail -8 $O/${T}_mica_cli_tests.log
timeout 600 python bench.py --workload mica > $O/${T}_mica_bench.log 2>&1 || exit 1
tail -1 $O/${T}_mica_bench.log > $O/${T}_mica_line.json; cut -c1-300 $O/${T}_mica_line.json; grep -o '"kernel_ms_per_step.*' $O/${T}_mica_line.json | cut -c1-400
if [ "$2" = "ncu" ]; then
  SS="python bench.py --workload mica --steps 1 --warmup 3 --no-cpu-baseline"
  k=k5_permutations
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/${T}_$k $SS > $O/${T}_ncu_$k.log 2>&1
  python tools/ncu_summary.py kernel $O/${T}_$k.ncu-rep $O/${T}_$k.txt
  ncu -i $O/${T}_$k.ncu-rep --page source --csv --print-source sass > /tmp/${T}_src.csv 2>/dev/null
  python tools/ncu_src.py /tmp/${T}_src.csv 14 >> $O/${T}_$k.txt 2>&1
  ncu -i $O/${T}_$k.ncu-rep --page details 2>/dev/null | grep -i -A3 "stall\|Warp Cycles Per Issued\|Executed Ipc\|Branch Efficiency\|Avg. Active Threads" | head -60 >> $O/${T}_$k.txt
  rm -f $O/${T}_$k.ncu-rep
  tail -50 $O/${T}_$k.txt
fi
