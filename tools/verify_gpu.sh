set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final_bench.log | cut -c1-600
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/final_bench_ref.log | cut -c1-400
timeout 300 python bench.py --workload mica > gpurun_out/final_mica.log 2>&1; echo "mica rc=$?"; tail -1 gpurun_out/final_mica.log | cut -c1-300; tail -1 gpurun_out/final_mica.log | grep -o '"roofline.*' | cut -c1-900
# compute-sanitizer is closed on this pool (tools/k5_sanitize.py is the small pass it would run over mica's kernels)
