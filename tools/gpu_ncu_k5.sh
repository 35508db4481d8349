cd $GRAFT_REPO_ROOT
O=gpurun_out
T=${1:-r2w}
SS="python bench.py --workload mica --steps 1 --warmup 3 --no-cpu-baseline"
k=k5_permutations
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $O/${T}_$k $SS > $O/${T}_ncu_$k.log 2>&1
python tools/ncu_summary.py kernel $O/${T}_$k.ncu-rep $O/${T}_$k.txt
ncu -i $O/${T}_$k.ncu-rep --page source --csv --print-source sass > /tmp/${T}_src.csv 2>/dev/null
python tools/ncu_src.py /tmp/${T}_src.csv 14 >> $O/${T}_$k.txt 2>&1
ncu -i $O/${T}_$k.ncu-rep --page details 2>/dev/null | grep -A3 "Warp Cycles Per Issued\|Executed Ipc\|Branch Efficiency\|Avg. Active Threads" | head -40 >> $O/${T}_$k.txt
rm -f $O/${T}_$k.ncu-rep
head -26 $O/${T}_$k.txt | cut -c1-110; grep "stall reasons\|Active Threads\|Ipc Active" $O/${T}_$k.txt
