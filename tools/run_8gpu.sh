#!/bin/bash
# One N-GPU box: training bench (both all-reduce routes), sharded inference (configs 2 and 4), gradient all-reduce
# check -> gpurun_out/
N=${1:-8}
TR="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r02_${N}gpu.json
CTXNERF_NCCL=0 $TR --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 --no-sustained 2>/dev/null | tail -1 > gpurun_out/bench_r02_${N}gpu_torch_allreduce.json
$TR --master-port 29512 bench.py --cfg4 --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/cfg4_${N}gpu.json
$TR --master-port 29513 bench.py --cfg2 --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/cfg2_${N}gpu.json
$TR --master-port 29514 tools/dist_check.py 2>&1 | grep "^rank" | sort > gpurun_out/dist_check_${N}gpu.txt
for f in bench_r02_${N}gpu bench_r02_${N}gpu_torch_allreduce; do
  python -c "import json,sys; d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce'), d.get('sustained'))"
done
cat gpurun_out/cfg4_${N}gpu.json gpurun_out/cfg2_${N}gpu.json gpurun_out/dist_check_${N}gpu.txt
