#!/bin/bash
# Short gpurun call: GPU tests, bench with CUDA-graph replay on and off, optional ncu capture.
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|graph vs eager" $OUT/pytest_gpu.log | tail -15
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_graph.json 2> $OUT/bench_graph.err; echo "bench(graph) exit $?"; tail -c 1500 $OUT/bench_graph.json; tail -5 $OUT/bench_graph.err
RORL_CUDA_GRAPH=0 timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_eager.json 2> $OUT/bench_eager.err; echo "bench(eager) exit $?"; tail -c 600 $OUT/bench_eager.json
if [ -n "$NCU_SEL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'selscan_(fwd|bwd)' -s 40 -c 4 -f -o $OUT/selscan \
    python tools/bench_kernels.py --only selscan > $OUT/ncu_selscan.log 2>&1; echo "ncu selscan exit $?"
fi
if [ -n "$EXTRA_CMD" ]; then bash -c "$EXTRA_CMD" > $OUT/extra.log 2>&1; echo "extra exit $?"; tail -30 $OUT/extra.log; fi
