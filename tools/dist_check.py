"""torchrun --nproc-per-node N tools/dist_check.py : the all-reduced gradient bucket of N ranks (each with
its own 4096-ray batch) equals the sum of the per-batch gradients computed by one process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
import torch.distributed as dist
from ctxnerf.train import NerfTrainer
from ctxnerf.workloads import orbit_camera

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
K, c2w = orbit_camera()
tr = NerfTrainer(800, 800, K, c2w, perturb=0.0, device=dev, seed=0)       # perturb=0: deterministic sampling
batches = []
for r in range(world):
    g = torch.Generator().manual_seed(100 + r)
    batches.append((torch.randint(0, 640000, (4096,), generator=g).to(dev), torch.rand(4096, 3, generator=g).to(dev)))
# distributed: my batch, one all-reduce
tr.step(*batches[rank], optimizer_step=False)
reduced = tr.bucket.grad.clone()
# single process reference on this rank: sum of the gradients of every batch
tr.world_size_backup = None
ref = torch.zeros_like(reduced)
for b in batches:
    tr.bucket.zero_grad()
    f = None
    import ctxnerf.dist as cd
    saved = cd.FlatBucket.all_reduce
    cd.FlatBucket.all_reduce = lambda self, group=None, async_op=False: None
    tr.step(*b, optimizer_step=False)
    cd.FlatBucket.all_reduce = saved
    ref += tr.bucket.grad
torch.cuda.synchronize()
err = (reduced - ref).abs().max().item() / ref.abs().max().item()
cos = torch.nn.functional.cosine_similarity(reduced, ref, dim=0).item()
print(f"rank {rank}/{world}: all-reduced vs single-process sum: max rel err {err:.3e}, cosine {cos:.8f}", flush=True)
assert err < 1e-3 and cos > 0.999999
dist.barrier()
dist.destroy_process_group()
