#!/bin/bash
# Final record of a round: GPU tests, smoke, both bench arms, ncu launch list of the bench command.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -1 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
gzip -f $OUT/launches.csv
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],2),"ms", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["clocks"], "launches", d["gpu_launches"])
r=json.loads(open("$OUT/bench_reference.json").read().strip().splitlines()[-1]); print("reference", round(r["value"],1))
PY
