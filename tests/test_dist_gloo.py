"""Host-side multi-GPU logic on CPU: world_size-2 gloo.  Ray sharding, per-rank
RNG streams, the flat gradient bucket and its single all-reduce (the only
collective of the path)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctxnerf.dist import FlatBucket, rank_generator, shard_rays, world as world_fn
        assert world_fn() == (rank, world)
        torch.manual_seed(0)                      # identical initial weights on every rank
        nets = [torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3)) for _ in range(2)]
        before = [p.detach().clone() for n in nets for p in n.parameters()]
        bucket = FlatBucket(nets)
        # parameters became views of one flat tensor, values preserved, grads are views of one bucket
        off = 0
        for p, b in zip(bucket.params, before):
            assert torch.equal(p.detach(), b)
            assert p.data_ptr() == bucket.flat.data_ptr() + 4 * off
            assert p.grad.data_ptr() == bucket.grad.data_ptr() + 4 * off
            off += p.numel()
        assert off == bucket.numel
        # this rank's shard of a global batch of 10 rays
        g = torch.Generator().manual_seed(123)
        x_all, y_all = torch.randn(10, 5, generator=g), torch.randn(10, 3, generator=g)
        lo, hi = shard_rays(10, rank, world)
        bucket.zero_grad()
        loss = sum(((n(x_all[lo:hi]) - y_all[lo:hi]) ** 2).sum() for n in nets)
        loss.backward()                            # autograd accumulates in place into the bucket views
        assert bucket.params[0].grad.data_ptr() == bucket.grad.data_ptr()
        bucket.all_reduce()
        # reference: single-process gradient of the concatenated batch
        torch.manual_seed(0)
        ref = [torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3)) for _ in range(2)]
        sum(((n(x_all) - y_all) ** 2).sum() for n in ref).backward()
        flat_ref = torch.cat([p.grad.reshape(-1) for n in ref for p in n.parameters()])
        torch.testing.assert_close(bucket.grad, flat_ref, rtol=1e-5, atol=1e-6)
        # per-rank RNG streams differ, are reproducible
        a = torch.randint(0, 640000, (8,), generator=rank_generator(0, rank))
        b = torch.randint(0, 640000, (8,), generator=rank_generator(0, rank))
        assert torch.equal(a, b)
        gathered = [torch.zeros_like(a) for _ in range(world)]
        dist.all_gather(gathered, a)
        assert not torch.equal(gathered[0], gathered[1])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


class _FakeTrainer:
    """Stands in for NerfTrainer in the host-logic test: 'renders' a deterministic function of the pixel id."""
    H, W = 5, 7
    device = torch.device("cpu")

    def render(self, ray_idx):
        i = (torch.arange(self.H * self.W) if ray_idx is None else ray_idx).float()
        return dict(rgb_map=torch.stack([i, 2 * i, 3 * i], -1), disp_map=i + 0.5, acc_map=i * 0.25, depth_map=-i,
                    rgb0=torch.stack([i, i, i], -1))

    def render_view(self, H, W, K, c2w, n_samples=192, sphere=None):
        base = float(c2w)                       # the fake "pose" is just a number identifying the view
        img = torch.full((H, W), base) + torch.arange(W).float()
        return dict(rgb_map=torch.stack([img, img + 1, img + 2], -1), acc_map=img * 0.5, depth_map=img * 2,
                    disp_map=img)


def _render_worker(rank, world, port, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctxnerf.dist import render_image_sharded, render_views_sharded, shard_rays
        tr = _FakeTrainer()
        # config 2: row blocks (35 pixels over 2 ranks: 18 + 17), local result and gathered image
        local = render_image_sharded(tr, gather=False)
        lo, hi = shard_rays(35, rank, world)
        assert local["rows"] == (lo, hi) and local["rgb_map"].shape == (hi - lo, 3)
        full = render_image_sharded(tr, gather=True)
        if rank == 0:
            ref = tr.render(None)
            for k in ("rgb_map", "disp_map", "acc_map", "depth_map", "rgb0"):
                assert torch.equal(full[k].reshape(ref[k].shape), ref[k]), k
            assert full["rgb_map"].shape == (5, 7, 3)
        # config 4: 5 views round-robin over 2 ranks (rank 0: views 0, 2, 4; rank 1: views 1, 3)
        cams = [(None, float(10 * v)) for v in range(5)]
        mine = render_views_sharded(tr, cams, 4, 6, gather=False)
        assert sorted(mine) == [v for v in range(5) if v % world == rank]
        allv = render_views_sharded(tr, cams, 4, 6, gather=True)
        if rank == 0:
            assert sorted(allv) == [0, 1, 2, 3, 4]
            for v in range(5):
                ref = tr.render_view(4, 6, None, float(10 * v))
                assert torch.equal(allv[v]["rgb_map"], ref["rgb_map"]) and torch.equal(allv[v]["depth_map"], ref["depth_map"])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_inference_world2_gloo():
    """render_image_sharded (config 2: row blocks) and render_views_sharded (config 4: views round-robin) with their
    gathers to rank 0, on two gloo ranks with a fake renderer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_render_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_rays_covers_everything():
    from ctxnerf.dist import shard_rays
    for n, w in ((640000, 8), (10, 3), (7, 8), (0, 2)):
        spans = [shard_rays(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_reference_arm_of_bench_runs_on_cpu():
    """bench.py --impl reference (the oracle port on the host cores) prints the contract's JSON line."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-rays", "256"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: calling an op on CPU tensors raises instead of silently computing."""
    from ctxnerf import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.CtxNerfError):
        ops.composite(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.zeros(2, 3))
    from ctxnerf import run_nerf_helpers as rh
    with pytest.raises(_lib.CtxNerfError):
        rh.sample_pdf(torch.zeros(2, 5), torch.zeros(2, 4), 8, det=True)
    from ctxnerf.texture import texture_mapping
    with pytest.raises(_lib.CtxNerfError):
        texture_mapping(torch.zeros(1, 4, 2), torch.zeros(1, 3, 8, 8), "bilinear")
    with pytest.raises(_lib.CtxNerfError):
        texture_mapping(torch.zeros(1, 4, 2), torch.zeros(1, 3, 8, 8), "bicubic")      # CPU tensors: no fallback
    with pytest.raises(_lib.CtxNerfError):
        texture_mapping(torch.zeros(1, 4, 2), torch.zeros(1, 3, 8, 8), "lanczos")
    with pytest.raises((_lib.CtxNerfError, RuntimeError)):
        rh.NeRF2D(D=8, W=256, input_ch=42, output_ch=3, skips=[4])(torch.zeros(4, 42))


def _comm_worker(rank, world, port, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # only rank 1 cannot find NCCL: the agreement step must make BOTH ranks give up before any rendezvous
    if rank == 1:
        os.environ["CTXNERF_NCCL_LIB"] = "/nonexistent/libnccl.so.2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ctxnerf._lib import CtxNerfError
        from ctxnerf.dist import BucketComm, dist_backend
        assert dist_backend() == "gloo"
        try:
            BucketComm(torch.device("cpu"))
            q.put((rank, "no error"))
        except CtxNerfError as e:
            assert "unavailable on at least one rank" in str(e)
            # the process group is still usable afterwards (nobody is stuck in a half-entered collective)
            t = torch.tensor([rank + 1])
            dist.all_reduce(t)
            q.put((rank, "agreed" if int(t) == 3 else "bad sum"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_communicator_availability_is_agreed_across_ranks_gloo():
    """BucketComm (the library's NCCL binding) when ONE rank cannot load NCCL: every rank raises the same error after
    the agreement all-reduce, none enters the token broadcast / the NCCL rendezvous alone, and the process group stays
    usable -- which is what lets NerfTrainer fall back to torch.distributed on all ranks together."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_comm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "agreed"), (1, "agreed")], res
