#!/bin/bash
# gpurun call for profiling only: bench line, torch-profiler table, ncu launch list, selected microbenchmarks.
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1200 $OUT/bench.json
timeout 300 python tools/profile_step.py > $OUT/profile_step.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
gzip -f $OUT/launches.csv
if [ -n "$KERNELS" ]; then timeout 600 python tools/bench_kernels.py --only $KERNELS --out $OUT/kernels.json > $OUT/kernels.log 2>&1; echo "kernels exit $?"; cat $OUT/kernels.log | cut -c1-200; fi
if [ -n "$EXTRA_CMD" ]; then bash -c "$EXTRA_CMD" > $OUT/extra.log 2>&1; echo "extra exit $?"; tail -30 $OUT/extra.log; fi
