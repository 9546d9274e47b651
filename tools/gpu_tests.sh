#!/bin/bash
# Short GPU call: (optionally) regenerate the GPU-side fixtures, then the parity tests selected by $K (pytest -k).
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$GOLDEN" ]; then
  timeout 600 python tests/golden/make_golden_gpu.py $OUT/golden_gpu > $OUT/golden_gpu.log 2>&1; echo "golden_gpu exit $?"; grep -E "wrote|FAILED|Error" $OUT/golden_gpu.log
  cp $OUT/golden_gpu/*.npz tests/golden/ 2>/dev/null
fi
timeout 1500 python -m pytest tests -m gpu -q -rs ${K:+-k "$K"} > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR|Error:|worst" $OUT/pytest_gpu.log | head -40
if [ -n "$EXTRA_CMD" ]; then bash -c "$EXTRA_CMD" > $OUT/extra.log 2>&1; echo "extra exit $?"; tail -40 $OUT/extra.log; fi
