set -x
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2q_ref_arm.log 2>&1; echo "ref rc=$?"; tail -5 gpurun_out/r2q_ref_arm.log | cut -c1-1500
( time timeout 900 python bench.py ) > gpurun_out/r2q_default.log 2>&1; echo "ours rc=$?"; tail -5 gpurun_out/r2q_default.log | cut -c1-300; grep -o '"cpu_baseline".*' gpurun_out/r2q_default.log | cut -c1-1500
