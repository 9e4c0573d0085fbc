"""Measure write-only / read-only / copy HBM bandwidth with torch fills and reductions (reference points for the
record-writing kernels)."""
import torch, json
dev = torch.device("cuda:0")
n = 1 << 30   # 4 GB of fp32
x = torch.empty(n, device=dev, dtype=torch.float32)
y = torch.empty(n, device=dev, dtype=torch.float32)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = {}
ms = t(lambda: x.zero_()); out["memset_write_GBs"] = 4 * n / ms / 1e6
ms = t(lambda: x.fill_(1.5)); out["fill_write_GBs"] = 4 * n / ms / 1e6
ms = t(lambda: y.copy_(x)); out["copy_rw_GBs"] = 8 * n / ms / 1e6
ms = t(lambda: x.sum()); out["sum_read_GBs"] = 4 * n / ms / 1e6
print(json.dumps(out))
