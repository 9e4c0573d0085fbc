#!/bin/bash
# One 8-GPU box: training bench, sharded inference (configs 2 and 4), gradient all-reduce check -> gpurun_out/
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r02_${N}gpu.json
$TR --master-port 29512 bench.py --cfg4 --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/cfg4_${N}gpu.json
$TR --master-port 29513 bench.py --cfg2 --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/cfg2_${N}gpu.json
$TR --master-port 29514 tools/dist_check.py 2>&1 | grep "all-reduced" > gpurun_out/dist_check_${N}gpu.txt
head -c 600 gpurun_out/bench_r02_${N}gpu.json; echo; cat gpurun_out/cfg4_${N}gpu.json gpurun_out/cfg2_${N}gpu.json gpurun_out/dist_check_${N}gpu.txt
