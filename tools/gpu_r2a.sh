#!/bin/bash
# Round-2 first GPU call: reference-generated cgpt fixtures, GPU parity tests, both bench arms, kernel table incl. the
# reference's own kernels on the same box.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python tests/golden/make_golden_gpu.py $OUT/golden_gpu > $OUT/golden_gpu.log 2>&1; echo "golden_gpu exit $?"; tail -5 $OUT/golden_gpu.log
cp $OUT/golden_gpu/*.npz tests/golden/ 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -rs -x --deselect tests/test_update_gpu.py::test_update_matches_reference > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
timeout 900 python -m pytest tests/test_update_gpu.py -m gpu -q -rs -k test_update_matches_reference > $OUT/pytest_update.log 2>&1; echo "pytest update exit $?"; tail -25 $OUT/pytest_update.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; tail -c 600 $OUT/bench_reference.json
timeout 900 python tools/bench_kernels.py --out $OUT/kernels.json > $OUT/kernels.log 2>&1; echo "kernels exit $?"; grep -v Warning $OUT/kernels.log | cut -c1-220
