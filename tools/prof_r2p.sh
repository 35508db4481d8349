# final evidence of round 2 (tag r2p): default bench line, launch list, ncu captures of the kernels changed since r2m
# (k2_tiles: conflict-free operand reads; k3_simulate: round keys from the launch parameters), the other two workloads
cd $GRAFT_REPO_ROOT
T=${1:-r2p}
O=gpurun_out
python bench.py > $O/${T}_bench_default.log 2>&1 || exit 1
tail -1 $O/${T}_bench_default.log > $O/${T}_bench_line_default.json
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv $S > $O/${T}_ncu_launch.log 2>&1
python tools/ncu_summary.py launches $O/${T}_launches.csv $O/${T}_launches_summary.txt
SS="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rep-cpu 128"
summ() {
  python tools/ncu_summary.py kernel $O/$1.ncu-rep $O/$1.txt
  ncu -i $O/$1.ncu-rep --page source --csv --print-source sass > /tmp/$1_src.csv 2>/dev/null
  python tools/ncu_src.py /tmp/$1_src.csv 12 >> $O/$1.txt 2>&1
  rm -f $O/$1.ncu-rep
}
for k in k2_tiles k3_simulate; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $O/${T}_$k $SS > $O/${T}_ncu_$k.log 2>&1
  summ ${T}_$k
done
python bench.py --workload proteins --steps 3 --warmup 3 > $O/${T}_proteins.log 2>&1; tail -1 $O/${T}_proteins.log > $O/${T}_proteins_line.json
python bench.py --workload clustering --steps 2 --warmup 1 > $O/${T}_clustering.log 2>&1; tail -1 $O/${T}_clustering.log > $O/${T}_clustering_line.json
rm -f $O/${T}_launches.csv
tail -c 400 $O/${T}_bench_line_default.json
