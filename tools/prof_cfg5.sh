# ncu evidence for configs[4] (20 000-site x 200-taxon protein clustering): bash tools/prof_cfg5.sh [tag]
cd $GRAFT_REPO_ROOT
T=${1:-r1k}
S="python tools/run_cfg5.py --null 1"
$S > gpurun_out/${T}_cfg5_line.json 2> gpurun_out/${T}_cfg5_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_cfg5_launches.csv $S > gpurun_out/${T}_cfg5_ncu_launch.log 2>&1
SS="python tools/run_cfg5.py --null 0"
for k in k1_down k1_up k2_tiles; do
  ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o gpurun_out/${T}_cfg5_$k $SS > gpurun_out/${T}_cfg5_ncu_$k.log 2>&1
done
cat gpurun_out/${T}_cfg5_line.json
