#!/bin/bash
# Round-2 measurement call: GPU parity tests, both bench arms, other encoder configurations, kernel table, eager profile.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 2500 $OUT/bench.json; tail -3 $OUT/bench.err
timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; tail -c 700 $OUT/bench_reference.json
for cfg in "td3 gilr" "td3 lru" "sac gru" "sac cgpt_h8_l6_p0.1_ml1024_rms"; do
  set -- $cfg
  RORL_BENCH_ALGO=$1 RORL_BENCH_ENCODER=$2 timeout 600 python bench.py --no-cpu-baseline --no-strong --steps 10 > $OUT/bench_$1_${2%%_*}.json 2>> $OUT/bench.err
  echo "$cfg: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_$1_${2%%_*}.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2),'ms', round(d['value']), 'steps/s  e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])")"
done
timeout 900 python tools/bench_kernels.py --out $OUT/kernels.json > $OUT/kernels.log 2>&1; echo "kernels exit $?"; grep -v Warning $OUT/kernels.log | cut -c1-200
RORL_PROFILE_LAUNCHES=1 timeout 600 python tools/profile_step.py > $OUT/profile.txt 2>&1; head -45 $OUT/profile.txt | grep -v Warn | cut -c1-120
