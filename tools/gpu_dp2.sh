#!/bin/bash
# 2-GPU data-parallel check: bench at N=2 (torchrun, NCCL) with CUDA-graph segments, then eager; bounded by timeouts.
TAG=${1:-dp2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "n2 exit $?"; tail -c 1100 $OUT/bench_n2.json; tail -3 $OUT/bench_n2.err
RORL_CUDA_GRAPH=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_n2_eager.json 2> $OUT/bench_n2_eager.err; echo "n2 eager exit $?"; tail -c 400 $OUT/bench_n2_eager.json
timeout 200 python bench.py --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "n1 exit $?"; tail -c 500 $OUT/bench_n1.json
if [ -n "$EXTRA_CMD" ]; then bash -c "$EXTRA_CMD" > $OUT/extra.log 2>&1; echo "extra exit $?"; tail -30 $OUT/extra.log | cut -c1-220; fi
