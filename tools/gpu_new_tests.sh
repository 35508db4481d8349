set -x
bash tools/ab.sh "CMB_X=0" "CMB_X=1"
python bench.py --workload clustering --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); print('clustering ms/step %.1f kernels %s' % (l['ms_per_step'], {k: round(v, 1) for k, v in l['kernel_ms_per_step'].items()}))"
