set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final_bench.log | cut -c1-600
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/final_bench_ref.log | cut -c1-400
