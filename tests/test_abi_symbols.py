"""The C-ABI library loads and exports every symbol include/ctxnerf.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ctxnerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctx_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for need in ("ctx_posenc_fwd", "ctx_posenc_bwd", "ctx_raygen_fwd", "ctx_stratified_fwd", "ctx_ndc_fwd",
                 "ctx_ndc_bwd", "ctx_composite_fwd", "ctx_composite_bwd", "ctx_resample_fwd", "ctx_resample_bwd",
                 "ctx_mlp_describe", "ctx_mlp_pack", "ctx_mlp_fwd"):
        assert need in syms


def test_integration_guide_maps_every_declared_symbol():
    """INTEGRATION.md names the reference interface each entry point replaces: no exported symbol without a row."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in declared_symbols() if s not in doc]
    assert not missing, missing


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_header(libpath):
    from ctxnerf import _lib
    assert set(declared_symbols()) <= set(_lib._SIGNATURES), set(declared_symbols()) - set(_lib._SIGNATURES)
    assert _lib.lib().ctx_abi_version() == 3
    assert b"bad argument" in _lib.lib().ctx_error_string(-1)


def test_argument_errors_do_not_need_a_gpu(libpath):
    """Negative sizes / null pointers are rejected before any launch."""
    from ctxnerf import _lib
    lib = _lib.lib()
    assert lib.ctx_composite_fwd(None, None, None, None, 4, 64, 0, None, None, None, None, None, None) == -1
    assert lib.ctx_composite_fwd(None, None, None, None, 0, 64, 0, None, None, None, None, None, None) == 0
    assert lib.ctx_posenc_fwd(None, None, -1, 3, 10, 1, 1, None) == -1
    assert lib.ctx_resample_fwd(None, 0, 0, None, 0, None, None, 1, 0, None, 8, 1, 16, None, None, None, 0, 0,
                                None, None) == -1


def test_render_args_struct_layout_matches_the_library(libpath):
    """The ctypes mirror of CtxRenderArgs has the size the library was compiled with, and the driver rejects an
    empty argument block without touching the GPU."""
    from ctxnerf import _lib
    lib = _lib.lib()
    assert lib.ctx_render_args_bytes() == ctypes.sizeof(_lib.CtxRenderArgs)
    assert lib.ctx_render_rays(None, None) == -1
    a = _lib.CtxRenderArgs()
    a.n_rays, a.n_samples = 0, 64
    assert lib.ctx_render_rays(ctypes.byref(a), None) == 0          # no rays: nothing to do
    a.n_rays = 16
    assert lib.ctx_render_rays(ctypes.byref(a), None) == -1         # no network / buffers


def test_communicator_binding_loads_nccl_at_run_time(libpath):
    """ctx_comm_*: no link-time dependency on NCCL (the library must load without it), a clear error before
    ctx_comm_load, and -- with the NCCL copy torch ships -- the 2.x version check.  No GPU work."""
    import subprocess
    deps = subprocess.run(["ldd", libpath], capture_output=True, text=True).stdout
    assert "nccl" not in deps
    from ctxnerf import _lib
    lib = _lib.lib()
    buf = ctypes.create_string_buffer(128)
    if lib.ctx_comm_version() == 0:                     # not loaded yet in this process
        assert lib.ctx_comm_unique_id(buf) == -3
        assert lib.ctx_allreduce(ctypes.c_void_p(1), ctypes.c_void_p(16), 4, None) == -3
        assert lib.ctx_comm_load(b"/nonexistent/libnccl.so.2") == -3
        assert b"dlopen failed" in lib.ctx_comm_last_error()
        assert b"NCCL" in lib.ctx_error_string(-3)
    import torch                                        # (maps its bundled NCCL where the platform has one)
    from ctxnerf.dist import _loaded_nccl_path
    import glob
    path = _loaded_nccl_path()
    if path is None:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib",
                                       "libnccl.so.2"))
        path = cands[0] if cands else None
    if path is None:
        pytest.skip("no NCCL shared object in this environment")
    assert lib.ctx_comm_load(path.encode()) == 0
    assert 20000 <= lib.ctx_comm_version() < 30000
    assert lib.ctx_comm_unique_id(None) == -1
    assert lib.ctx_comm_init(None, 2, buf, 0) == -1
    handle = ctypes.c_void_p()
    assert lib.ctx_comm_init(ctypes.c_void_p(ctypes.addressof(handle)), 2, buf, 2) == -1      # rank out of range
    assert lib.ctx_allreduce(None, None, 0, None) == -1
    assert lib.ctx_comm_destroy(None) == 0


def test_mlp_describe_matches_the_reference_layer_structure(libpath):
    """NeRF2D(D=8, W=256, skips=[4]): layer 5 takes [x | h] (reference :81-83)."""
    from ctxnerf.mlp import NetDesc
    d = NetDesc(8, [4], 63, 0, 4)
    assert d.n_layers == 8
    # forward stream: L0 64x256, 7 hidden 256x256 (+64 on the skip layer), bf16
    # (+ one 16-wide bias chunk for each of the 6 layers that do not read the x buffer)
    assert d.w_bytes == 2 * 256 * (64 + 7 * 256 + 64) + 6 * 2 * 256 * 16
    v = NetDesc(8, [4], 63, 27, 4)
    assert v.n_layers == 10
    assert v.w_bytes == d.w_bytes + 2 * (256 * 256 + 128 * (256 + 32)) + 2 * 256 * 16


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/ctxnerf.h has as many parameters, and pointers / 64-bit integers in the same
    positions, as the ctypes signature the Python host binds it with."""
    from ctxnerf import _lib
    text = open(os.path.join(ROOT, "include", "ctxnerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(?:int|const char\*)\s+(ctx_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 20
    for name, args in protos:
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        _, argtypes = _lib._SIGNATURES[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for p, t in zip(params, argtypes):
            if "*" in p:
                assert t is ctypes.c_void_p or t is ctypes.c_char_p, (name, p, t)
            elif p.startswith("int64_t"):
                assert t is ctypes.c_int64, (name, p, t)
            elif p.startswith("uint64_t"):
                assert t is ctypes.c_uint64, (name, p, t)
            elif p.startswith("uint32_t"):
                assert t in (ctypes.c_uint32, ctypes.c_int), (name, p, t)
            elif p.startswith("float"):
                assert t is ctypes.c_float, (name, p, t)
            elif p.startswith("int "):
                assert t is ctypes.c_int, (name, p, t)


def test_diagnostics_are_not_part_of_the_product_library(libpath):
    """tc_selftest micro-benchmarks and the ctx_mlp_set_* profiling hooks live in libctxnerf_diag.so
    (include/ctxnerf_diag.h); the product library and its header carry none of them."""
    from ctxnerf import _lib
    prod = ctypes.CDLL(libpath)
    diag = ctypes.CDLL(_lib.DIAG_LIB_PATH)
    text = open(os.path.join(ROOT, "include", "ctxnerf_diag.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    diag_syms = sorted(set(re.findall(r"\b(ctx_[a-z0-9_]+)\s*\(", text)))
    assert "ctx_tcgen05_selftest" in diag_syms and "ctx_mlp_set_prof_buffer" in diag_syms
    for sym in diag_syms:
        assert hasattr(diag, sym), sym
        assert not hasattr(prod, sym), f"{sym} leaked into the product library"
        assert sym not in declared_symbols()
    for sym in declared_symbols():          # the diagnostics build is a superset
        assert hasattr(diag, sym), sym
