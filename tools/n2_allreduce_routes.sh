set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist2_ctx.log 2>&1; echo rc=$?
CTXNERF_NCCL=0 timeout 300 $TR --master-port 29512 tools/dist_check.py > gpurun_out/dist2_torch.log 2>&1; echo rc=$?
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained > gpurun_out/bench2_ctx.log 2>&1; echo rc=$?
CTXNERF_NCCL=0 timeout 400 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained > gpurun_out/bench2_torch.log 2>&1; echo rc=$?
timeout 400 $TR --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --no-sustained > gpurun_out/bench2_ctx_b.log 2>&1; echo rc=$?
grep -h "rank" gpurun_out/dist2_*.log | grep -v Warn
for f in gpurun_out/bench2_*.log; do tail -1 $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('allreduce'))"; done
