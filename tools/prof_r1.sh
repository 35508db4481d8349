# ncu evidence for profiles/ (run on the GPU box: bash tools/prof_r1.sh)
cd $GRAFT_REPO_ROOT
S="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$S > gpurun_out/r1f_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1f_launches.csv $S > gpurun_out/r1f_ncu_launch.log 2>&1
SS="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rep-cpu 128"
$SS > gpurun_out/r1f_plain_small.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k1_up -s 4 -c 1 -o gpurun_out/r1f_k1_up $SS > gpurun_out/r1f_ncu_up.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_down -s 4 -c 1 -o gpurun_out/r1f_k1_down $SS > gpurun_out/r1f_ncu_down.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2_tiles -s 2 -c 1 -o gpurun_out/r1f_k2_tiles $SS > gpurun_out/r1f_ncu_tiles.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k3_simulate -s 4 -c 1 -o gpurun_out/r1f_k3_sim $SS > gpurun_out/r1f_ncu_sim.log 2>&1
tail -c 300 gpurun_out/r1f_plain.log
