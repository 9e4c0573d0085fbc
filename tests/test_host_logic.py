"""Host-side contracts that need no GPU: the product never routes through the oracle or any CPU path, fails loudly
without its CUDA library or device, and the ray sharding of the multi-GPU path partitions exactly."""
import ast
import os
import subprocess
import sys

import pytest
import torch
from hypothesis import given, settings, strategies as st

from conftest import PKG, ROOT


def _imports(path):
    tree = ast.parse(open(path).read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            out += [(a.name, node.lineno) for a in node.names]
        elif isinstance(node, ast.ImportFrom) and node.module:
            out.append((node.module, node.lineno))
    return out


def test_package_never_imports_the_oracle_or_the_reference():
    """oracle/ is test infrastructure: nothing under contexture-nerf_b200/ may import it (or the reference tree)."""
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                for mod, line in _imports(os.path.join(dirpath, f)):
                    if mod.split(".")[0] in ("oracle", "src") or "reference" in mod:
                        bad.append((f, line, mod))
    assert not bad, bad


def test_bench_uses_the_oracle_only_in_its_cpu_legs():
    """bench.py may execute oracle/ in cpu_baseline / --impl reference / --cfg1 only: every import of it sits inside
    one of those functions, none at module level and none in the timed GPU path."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    allowed = {"cpu_step_sample", "run_reference", "run_cfg1"}
    for node in tree.body:                                  # module level: no oracle import
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            names = [a.name for a in node.names] + [getattr(node, "module", "") or ""]
            assert not any(n.split(".")[0] == "oracle" for n in names)
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, (ast.Import, ast.ImportFrom)) and
                   (("oracle" in (getattr(n, "module", "") or "")) or any(a.name.split(".")[0] == "oracle" for a in n.names))
                   for n in ast.walk(fn))
        if uses:
            assert fn.name in allowed, fn.name


def test_no_cpu_fallback_every_entry_point_raises_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("this is the CPU-only contract")
    from ctxnerf import _lib, ops, run_nerf_helpers as rh
    calls = [lambda: ops.posenc(torch.rand(4, 3), 10),
             lambda: rh.raw2outputs(torch.rand(2, 8, 4), torch.rand(2, 8), torch.rand(2, 3)),
             lambda: rh.sample_pdf(torch.rand(2, 8), torch.rand(2, 7), 4, det=True),
             lambda: rh.get_rays(4, 4, [[1., 0, 2], [0, 1., 2], [0, 0, 1]], torch.eye(4)[:3]),
             lambda: rh.get_embedder(10)[0](torch.rand(3, 2)),
             lambda: rh.NeRF2D(input_ch=42, output_ch=3)(torch.rand(5, 42))]
    for c in calls:
        with pytest.raises(_lib.CtxNerfError, match="no CPU fallback"):
            c()


def test_missing_library_is_a_loud_error():
    """A process that cannot find libctxnerf.so must say so (and how to build it), not fall back to anything."""
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "from ctxnerf import _lib\n"
            "_lib.LIB_PATH = '/nonexistent/libctxnerf.so'; _lib._lib = None\n"
            "try:\n    _lib.lib()\nexcept _lib.CtxNerfError as e:\n    print('LOUD', 'build' in str(e) and 'fallback' in str(e))\n" % (ROOT, PKG))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "LOUD True" in out.stdout, out.stdout + out.stderr


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 5000), world=st.integers(1, 16))
def test_shard_rays_partitions_exactly(n, world):
    from ctxnerf.dist import shard_rays
    blocks = [shard_rays(n, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (a, b), (c, d) in zip(blocks, blocks[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
