#!/bin/bash
# gpurun call for profiling: bench line, ncu launch list of the bench command, ncu --set full of chosen kernels,
# bench lines of the other encoder configurations.
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1500 $OUT/bench.json
timeout 300 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
for cfg in "td3 gilr" "td3 lru" "sac gru" "sac mamba_s32_c16" "sac cgpt_h8_l6_p0.1_ml1024_rms"; do
  set -- $cfg
  RORL_BENCH_ALGO=$1 RORL_BENCH_ENCODER=$2 timeout 300 python bench.py --no-cpu-baseline --steps 10 > $OUT/bench_$1_${2%%_*}.json 2> $OUT/bench_$1_${2%%_*}.err; echo "$cfg exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$1_${2%%_*}.json").read().strip().splitlines()[-1]); print("$cfg", round(d["ms_per_step"],2), "ms", round(d["value"]), "steps/s, e2e", round(d["e2e"]["value"]))
except Exception as e: print("$cfg failed", e)
PY
done
kill $SMI
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
gzip -f $OUT/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 30 -c 3 -f -o $OUT/gemm \
    python tools/bench_kernels.py --only gemm > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
if [ -n "$EXTRA_CMD" ]; then bash -c "$EXTRA_CMD" > $OUT/extra.log 2>&1; echo "extra exit $?"; tail -30 $OUT/extra.log; fi
