# ncu evidence for profiles/ (run on the GPU box: bash tools/prof_r1.sh [tag])
cd $GRAFT_REPO_ROOT
T=${1:-r1j}
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$S > gpurun_out/${T}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv $S > gpurun_out/${T}_ncu_launch.log 2>&1
SS="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rep-cpu 128"
$SS > gpurun_out/${T}_plain_small.log 2>&1 || exit 1
for k in k1_up_mma k1_down_mma k2_tiles k2_paired k2_pvalues k3_simulate; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/${T}_$k $SS > gpurun_out/${T}_ncu_$k.log 2>&1
done
tail -c 300 gpurun_out/${T}_plain.log
