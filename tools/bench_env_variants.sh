run() { env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), round(d['step_tensor_frac'],4), ' '.join(k[4:]+' '+str(round(v['ms_per_launch'],3)) for k,v in d['kernels'].items()))
"; }
run X=0
run CTXNERF_FWD_TMA=1
run CTXNERF_LIB=$PWD/contexture-nerf_b200/ctxnerf/variants/libctxnerf_fwdpipe.so
run CTXNERF_LIB=$PWD/contexture-nerf_b200/ctxnerf/variants/libctxnerf_fwdpipe.so CTXNERF_FWD_TMA=1
