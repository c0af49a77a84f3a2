"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/ltu_b200.h declares; argument errors are reported without touching a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lintransunet_b200 import build, _native
    build.build(force=False)
    return _native.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ltu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ltu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from lintransunet_b200 import _native
    syms = declared_symbols()
    assert len(syms) >= 25
    assert set(syms) == set(_native.SIGNATURES), set(syms) ^ set(_native.SIGNATURES)
    for s in syms:
        assert getattr(lib, s) is not None


def test_sass_is_sm100a_only():
    import subprocess
    from lintransunet_b200 import _native
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_argument_errors_do_not_need_a_gpu(lib):
    assert lib.ltu_version() >= 100
    assert lib.ltu_gelu(None, 0, 0, None) == -1
    assert b"gelu" in lib.ltu_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.ltu_add_layernorm(p, p, p, p, p, 4, 100, 1e-6, 0, None) == -1      # C not in {128,256}
    assert b"C must be 128 or 256" in lib.ltu_last_error()
    assert lib.ltu_kv_reduce(p, p, 96, p, p, 0, 1, 10, 3, 0, None) == -1          # heads = 3
    assert lib.ltu_conv3d(p, 16, None, 0, 1, 4, 4, 4, 0, 5, 1, 1, 1, 2, p, None, 16, p, 0, 4, 4, 4, None, 0, None) == -1
    assert lib.ltu_head_d2s_softmax(p, p, None, None, 1, 2, 2, 2, 9, None) == -1   # dim_output > 8
    assert lib.ltu_kv_reduce_workspace(0, 0, 4) == 0
    assert lib.ltu_vote_decide(p, p, 3, 64, 2, 0.5, None) == -1                    # mode not in {0,1}
    assert b"vote_decide" in lib.ltu_last_error()
    assert lib.ltu_keep_largest_component_workspace(1000) == 16 + 8000
    assert lib.ltu_keep_largest_component(p, 3, 0b110, 2, 2, 2, 4, p, 1 << 20, None) == -1      # connectivity 4
    assert lib.ltu_keep_largest_component(p, 3, 0b1000, 2, 2, 2, 3, p, 1 << 20, None) == -1     # label 3 of 3 channels
    assert lib.ltu_keep_largest_component(p, 3, 0b110, 2, 2, 2, 3, p, 8, None) == -1            # workspace too small
    assert b"workspace" in lib.ltu_last_error()
    assert lib.ltu_overlap_counts(p, None, 3, 2, 2, 2, p, None) == -1
    # backward entry points
    assert lib.ltu_attn_bwd_workspace(1, 100, 3) == 0                              # heads = 3
    assert lib.ltu_attn_bwd(p, 128, p, p, 128, p, 128, p, p, p, p, 384, p, p, p, 0, 1, 100, 4, 0, None) == -1   # workspace 0
    assert b"workspace" in lib.ltu_last_error()
    assert lib.ltu_add_layernorm_bwd(p, p, p, p, p, p, p, p, 1 << 20, 4, 100, 1e-6, 0, None) == -1              # C = 100
    assert lib.ltu_gelu_bwd(p, p, p, 7, 0, None) == -1                             # n not a multiple of 4
    assert lib.ltu_instnorm_bwd(p, p, p, p, p, 1 << 20, 1, 64, 24, 1, 0, None) == -1   # C = 24
    assert lib.ltu_conv3d_wgrad(p, p, p, p, 1 << 20, 1, 4, 4, 4, 12, 4, 4, 4, 16, 3, 1, 1, 1, 1, 0, None) == -1  # Cin = 12
    assert lib.ltu_conv3d_wgrad(p, p, p, p, 1 << 20, 1, 4, 4, 4, 16, 3, 4, 4, 16, 3, 1, 1, 1, 1, 0, None) == -1  # Ho mismatch
    assert b"output size" in lib.ltu_last_error()
    assert lib.ltu_upsample_trilinear_bwd(p, p, 1, 2, 2, 2, 16, 3, 0, None) == -1  # depth factor 3
    assert lib.ltu_gate_bwd_workspace(1, 64, 24) > 0 and lib.ltu_gate_bwd_workspace(1, 64, 6) == 0
    assert lib.ltu_roi_resample_bwd(p, p, p, p, 1 << 20, 1, 8, 8, 4, 16, 10, 6, 9, 5, 0, 0, None) == -1         # eval <= roi
    assert lib.ltu_zero_insert(p, p, 1, 4, 4, 4, 16, 6, 7, 7, 2, 2, 2, 0, None) == -1                           # target too small


def test_product_refuses_cpu_tensors(lib):
    import torch
    from lintransunet_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gelu_(torch.zeros(8))
    with pytest.raises(RuntimeError):
        ops.kv_reduce(torch.zeros(1, 4, 128), torch.zeros(1, 4, 128), 4)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from lintransunet_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_native.NativeLibraryMissing):
        _native.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lintransunet_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|__import__\(.oracle", src, re.M), fn
