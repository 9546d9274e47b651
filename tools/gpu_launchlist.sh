#!/bin/bash
# ncu launch list (device time per launch) of two timed bench updates -> gpurun_out/<tag>/launches.csv.gz + bench line
TAG=${1:-ll}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py --no-cpu-baseline ${BENCH_ARGS} > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 400 $OUT/bench.json | head -c 400; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
gzip -f $OUT/launches.csv
