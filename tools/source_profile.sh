#!/bin/bash
# Per-source-line instruction counts of the issue-bound HBM kernels (ncu --import-source, source page exported on the
# box: the .ncu-rep is too large to bring back).  usage: tools/source_profile.sh <kernel regex> <tag> [launch-skip]
K=${1:-resample_fast}; TAG=${2:-resample}; SKIP=${3:-0}
mkdir -p gpurun_out/src
python tools/hbm_once.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c 1 -o gpurun_out/src/$TAG python tools/hbm_once.py > gpurun_out/src/$TAG.log 2>&1
ncu -i gpurun_out/src/$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src/${TAG}_source.csv 2> gpurun_out/src/${TAG}_source.err \
  || ncu -i gpurun_out/src/$TAG.ncu-rep --page source --csv > gpurun_out/src/${TAG}_source.csv 2>> gpurun_out/src/${TAG}_source.err
rm -f gpurun_out/src/$TAG.ncu-rep
ls -la gpurun_out/src
