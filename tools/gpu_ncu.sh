#!/bin/bash
# ncu --set full capture of a few launches:  NCU_K=<kernel regex> NCU_CMD="python tools/..." [NCU_S=skip] [NCU_C=count] bash tools/gpu_ncu.sh tag
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-0} -c ${NCU_C:-6} -f -o $OUT/cap \
    $NCU_CMD > $OUT/ncu.log 2>&1; echo "ncu exit $?"; tail -3 $OUT/ncu.log
ls -la $OUT
