"""CPU-side checks of the C-ABI library: it loads, and exports every symbol include/ctk.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import ctk
    if not os.path.exists(ctk.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ctk_build", os.path.join(ROOT, "torch-unet_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return ctk.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ctk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    import ctk
    declared = _declared_symbols()
    assert declared, "no declarations found in include/ctk.h"
    assert sorted(ctk.EXPORTED_SYMBOLS) == declared


def test_every_declared_symbol_is_exported(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_status_strings_and_no_compute_paths(lib):
    assert lib.ctk_abi_version() == 2
    assert lib.ctk_status_string(0) == b"ok"
    assert b"argument" in lib.ctk_status_string(-1)
    assert lib.ctk_pearson_workspace_bytes(0) == 0
    assert lib.ctk_pearson_workspace_bytes(256) == 256 * 8 * 8 * 8
    # argument validation happens before any CUDA call: null pointers are refused without a device
    assert lib.ctk_pearson_f32(None, 4, 65536, None, None, 0, None) == -1
    assert lib.ctk_conv3x3_tc_eval(None, 1, 16, 16, 64, None, 128, None, None, 0.01, None, 128, 0, 0, None) == -1
    assert lib.ctk_tile_ssim_workspace_bytes(0) == 0 and lib.ctk_tile_ssim_workspace_bytes(256) == 256 * (8 * 8 + 4) + 8
    assert lib.ctk_tile_ssim_f32(None, 4, 256, 256, None, None, 0, None) == -1           # null pointers
    assert lib.ctk_tile_ssim_f32(None, 0, 256, 256, None, None, 0, None) == 0            # empty batch: nothing to do
    assert lib.ctk_gemm_bf16_splitk(None, None, 128, 128, 64, 1, None, None) == -1
    # the deterministic two-stage reductions take caller-owned workspaces: sizes are queries, null pointers are refused
    assert lib.ctk_bn_bwd_reduce_workspace_bytes(0) == 0 and lib.ctk_bn_bwd_reduce_workspace_bytes(64) % (2 * 64 * 4) == 0
    assert lib.ctk_conv3x3_tc_raw_workspace_bytes(256) % (2 * 256 * 4) == 0
    assert lib.ctk_conv3x3_wgrad_tc_workspace_bytes(64, 128) >= 3 * 12 * 32 * 128 * 4
    assert lib.ctk_first_patch_gram_workspace_bytes(2) % ((18 + 171) * 8) == 0
    assert lib.ctk_conv3x3_wgrad_tc(None, None, 1, 16, 16, 64, 128, None, None, 0, None) == -1
    assert lib.ctk_bn_bwd_reduce_guarded(None, 1, 16, 16, None, None, None, None, None, 64, 0, None, 64, 0, 64, None, None,
                                         0.01, None, None, 0, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import ctk._lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.CtkError):
        L.load()
