# A/B of an environment switch on the bench: tools/ab_env.sh VAR "cfg algo encoder" ...   (VAR=1 vs VAR=0)
VAR=$1; shift
for cfg in "$@"; do
  set -- $cfg
  for v in 1 0; do
    env $VAR=$v RORL_BENCH_ALGO=$1 RORL_BENCH_ENCODER=$2 timeout 600 python bench.py --no-cpu-baseline --no-strong --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg $VAR=$v', round(d['ms_per_step'],3),'ms', round(d['value']), 'launches/step', d['gpu_launches']//(d['steps']+d['warmup']))"
  done
done
