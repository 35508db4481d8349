# second part of the r2 evidence: the wide (128-site) A = 4 mapping kernels and the protein tensor-core kernels
cd $GRAFT_REPO_ROOT
T=${1:-r2m}
O=gpurun_out
summ() {
  python tools/ncu_summary.py kernel $O/$1.ncu-rep $O/$1.txt
  ncu -i $O/$1.ncu-rep --page source --csv --print-source sass > /tmp/$1_src.csv 2>/dev/null
  python tools/ncu_src.py /tmp/$1_src.csv 12 >> $O/$1.txt 2>&1
  rm -f $O/$1.ncu-rep
}
SS="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rep-cpu 128"
for k in k1_up_mma k1_down_mma; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/${T}_$k $SS > $O/${T}_ncu_$k.log 2>&1
  summ ${T}_$k
done
SP="python bench.py --workload proteins --steps 1 --warmup 3"
$SP > $O/${T}_proteins_line.json 2>&1
for k in k1_up_mma20 k1_down_mma20; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o $O/${T}_$k $SP > $O/${T}_ncu_$k.log 2>&1
  summ ${T}_$k
done
python bench.py --workload clustering --steps 2 --warmup 1 > $O/${T}_clustering_line.json 2>&1
