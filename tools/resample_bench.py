"""sample_pdf / fused resample+merge throughput at 2^20 rays (fraction of the measured HBM copy bandwidth)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import ops
dev = torch.device("cuda:0")
HBM = 6540.2e9
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(n)]; b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e-3
R = 1 << 20
for S in (64, 128, 256):
    N = 2 * S
    z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]; w = torch.rand(R, S, device=dev)
    bins = z[:, :S - 1].contiguous(); wc = w[:, :S - 2].contiguous()
    by = R * (4 * (S - 1) + 4 * (S - 2) + 4 * N); bym = R * (4 * S + 4 * S + 4 * N + 4 * (S + N))
    with torch.no_grad():
        for rep in range(2):
            t2 = timeit(lambda: ops.resample(bins, wc, N, det=False, seed=3))
            t = timeit(lambda: ops.resample(bins, wc, N, det=True))
            t4 = timeit(lambda: ops.resample_merge(z, w, N, det=False, seed=3))
            t3 = timeit(lambda: ops.resample_merge(z, w, N, det=True))
            print(f"S={S} N={N}: sample_pdf det {by/t/HBM:.3f} rand {by/t2/HBM:.3f} | +merge det {bym/t3/HBM:.3f} rand {bym/t4/HBM:.3f}"
                  f"  ({t*1e3:.2f} {t2*1e3:.2f} {t3*1e3:.2f} {t4*1e3:.2f} ms)")
