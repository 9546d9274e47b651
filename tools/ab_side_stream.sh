# A/B of the side-stream overlap of independent context encoders (RORL_SIDE_STREAM=0 disables it)
for cfg in "sac smamba_s32_c16_b2_nln" "sac gru" "td3 gilr" "td3 lru"; do
  set -- $cfg
  for ss in 1 0; do
    RORL_SIDE_STREAM=$ss RORL_BENCH_ALGO=$1 RORL_BENCH_ENCODER=$2 timeout 600 python bench.py --no-cpu-baseline --no-strong --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg side=$ss', round(d['ms_per_step'],3),'ms', round(d['value']))"
  done
done
