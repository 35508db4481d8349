# A/B of environment switches on the GPU box: bash tools/ab.sh "VAR=a VAR2=b" "VAR=c" ...   (one bench line per setting)
cd $GRAFT_REPO_ROOT
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('ms/step %.2f  e2e %.2f  kernels %s' % (l['ms_per_step'], l['e2e']['ms_per_step'], {k: round(v, 2) for k, v in l['kernel_ms_per_step'].items()}))"
done
