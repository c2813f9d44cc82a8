"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/b200clip.h declares; the ctypes table matches the header's argument counts; ops refuse CPU tensors loudly."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "b200clip.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(b200clip_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def lib_path():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200_build", os.path.join(ROOT, "clip-for-dl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(verbose=False)


def test_library_exports_every_declared_symbol(lib_path):
    decl = _header_functions()
    assert len(decl) >= 25
    lib = ctypes.CDLL(lib_path)
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/b200clip.h but not exported"
    lib.b200clip_version.restype = ctypes.c_int
    assert lib.b200clip_version() >= 100


def test_ctypes_table_matches_header(lib_path):
    from b200clip import _lib
    decl = _header_functions()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, n in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name
    _lib.load()


def test_workspace_queries_are_host_only(lib_path):
    from b200clip import _lib
    lib = _lib.load()
    assert lib.b200clip_infonce_workspace_bytes(4096, 32768) > 0
    assert lib.b200clip_proj_bwd_workspace_bytes(4096, 2048, 512) >= 4096 * 512 * 8
    assert lib.b200clip_smallc_workspace_bytes(1000, 16, 512) > 0


def test_no_cpu_fallback():
    import b200clip
    x = torch.randn(8, 512)
    with pytest.raises(RuntimeError, match="CUDA"):
        b200clip.contrastive_loss(x, x, 0.07)
    with pytest.raises(RuntimeError, match="CUDA"):
        b200clip.ImageProjection(64, 512).eval()(torch.randn(4, 64))


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "clip-for-dl_b200", "b200clip")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "ref_head" not in src and "oracle" not in src.replace("# oracle", ""), fn
