#!/bin/bash
# Per-source-line instruction counts / stall samples of one MLP kernel launch (training-mode forward, dgrad or wgrad).
# usage: tools/source_profile_mlp.sh <kernel regex> <tag> [launch-skip]
K=${1:-mlp_fwd}; TAG=${2:-mlp_fwd}; SKIP=${3:-1}
mkdir -p gpurun_out/src
python tools/fwd_train_only.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c 1 -o gpurun_out/src/$TAG python tools/fwd_train_only.py > gpurun_out/src/$TAG.log 2>&1
ncu -i gpurun_out/src/$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src/${TAG}_source.csv 2> gpurun_out/src/${TAG}_source.err \
  || ncu -i gpurun_out/src/$TAG.ncu-rep --page source --csv > gpurun_out/src/${TAG}_source.csv 2>> gpurun_out/src/${TAG}_source.err
rm -f gpurun_out/src/$TAG.ncu-rep
ls -la gpurun_out/src
