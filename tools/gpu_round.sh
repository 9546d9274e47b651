#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), per-kernel microbench, ncu launch list of the bench
# command and ncu --set full captures of the scan kernels.  Everything lands in gpurun_out/<tag>/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r01b'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 3000 $OUT/bench.json
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; cat $OUT/bench_reference.json
timeout 600 python tools/bench_kernels.py --out $OUT/kernels.json > $OUT/kernels.log 2>&1; echo "kernels exit $?"
timeout 300 python tools/profile_step.py > $OUT/profile_step.txt 2>&1
kill $SMI
# launch list of the bench command (serialised, cold-cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
gzip -f $OUT/launches.csv
# full capture of the scan kernels at the workload shape
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'selscan_(fwd|bwd)' -s 40 -c 4 -f -o $OUT/selscan \
    python tools/bench_kernels.py --only selscan > $OUT/ncu_selscan.log 2>&1; echo "ncu selscan exit $?"
if [ -n "$NCU_EXTRA" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_EXTRA" -s 6 -c 6 -f -o $OUT/extra \
      python tools/bench_kernels.py --only ${NCU_EXTRA_ONLY:-gemm} > $OUT/ncu_extra.log 2>&1; echo "ncu extra exit $?"
fi
ls -la $OUT
