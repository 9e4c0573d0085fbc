TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29550
for w in 5 30 5 30; do
  $TR --master-port $P bench.py --gpus 2 --steps 20 --warmup $w --no-sustained 2>/dev/null | tail -1 > gpurun_out/w$w.json; P=$((P+1))
  python -c "import json; d=json.load(open('gpurun_out/w$w.json')); print('warmup', $w, 'value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), d['clocks'])"
done
