"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/specforge_b200.h declares,
and its pure-host entry points (layout / workspace planning / argument validation) behave."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from specforge_b200._lib import LIB_PATH, build_library, lib
    if not os.path.exists(LIB_PATH):
        build_library()
    return lib()


def test_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "specforge_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(sf_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18, names
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_param_layout_and_workspace(L):
    from specforge_b200.engine import P_COUNT, SfConfig, _declare
    from ctypes import c_int64
    _declare()
    # Qwen3-8B EAGLE3 draft (configs/qwen3-8b-eagle3.json): 399.52 M trainable parameters
    cfg = SfConfig(8, 2048, 7, 4096, 4096, 12288, 32, 8, 128, 151936, 32000, 0, 1, 40980, 1e-6, 0.8, 0, 1.0, 1.0)
    offs, sizes, total = (c_int64 * P_COUNT)(), (c_int64 * P_COUNT)(), c_int64()
    assert L.sf_eagle3_param_layout(cfg, offs, sizes, ctypes.byref(total)) == 0
    assert total.value == 399_523_840 + 0 or abs(total.value - 399.52e6) < 0.01e6
    # q,k,v adjacent and gate,up adjacent (fused GEMM operands)
    assert offs[2] == offs[1] + sizes[1] and offs[3] == offs[2] + sizes[2]
    assert offs[6] == offs[5] + sizes[5]
    ws = L.sf_eagle3_workspace_bytes(cfg)
    assert 20e9 < ws < 80e9, ws
    bad = SfConfig(8, 2048, 7, 4096, 4096, 12288, 32, 8, 96, 151936, 32000, 0, 1, 40980, 1e-6, 0.8, 0, 1.0, 1.0)
    assert L.sf_eagle3_workspace_bytes(bad) == 0
    assert b"head_dim" in L.sf_last_error()


def test_gemm_rejects_bad_arguments_without_a_gpu(L):
    from specforge_b200 import _lib
    rc = L.sf_gemm_bf16(None, 8, 0, None, 8, 0, None, 8, None, 0, 0, 16, 16, 0, 0, None)
    assert rc != 0 and b"empty" in L.sf_last_error()
    with pytest.raises(_lib.SfError):
        _lib.check(rc, "sf_gemm_bf16")


def test_dflash_param_layout_and_workspace(L):
    """sf_dflash_param_layout on the BASELINE config-4 draft (configs/qwen3-8b-dflash.json): names/sizes of DFlashDraftModel,
    q|k|v and gate|up adjacent (fused GEMM operands), workspace within the 180 GB part."""
    import ctypes
    from ctypes import c_int64
    from specforge_b200.dflash import GLOBALS, PER_LAYER, SfDflashConfig, _declare
    from oracle import dflash_oracle as D
    _declare(L)
    cfg = SfDflashConfig(4, 2048, 512, 16, 4096, 5, 12288, 32, 8, 128, 5, 151936, 151669, 41000, 1e-6, 0.0)
    n = L.sf_dflash_num_params(cfg)
    assert n == 11 * 5 + 3
    offs, sizes, total = (c_int64 * n)(), (c_int64 * n)(), c_int64()
    assert L.sf_dflash_param_layout(cfg, offs, sizes, ctypes.byref(total)) == 0
    names = [f"layers.{l}.{p}" for l in range(5) for p in PER_LAYER] + list(GLOBALS)
    shapes = D.param_shapes(D.DFlashConfig(hidden_size=4096, intermediate_size=12288, num_heads=32, num_kv_heads=8, head_dim=128,
                                           num_layers=5, num_target_feats=5, vocab_size=151936, block_size=16))
    assert set(names) == set(shapes)
    for i, nm in enumerate(names):
        numel = 1
        for s in shapes[nm]:
            numel *= s
        assert sizes[i] == numel, nm
        assert offs[i] == (0 if i == 0 else offs[i - 1] + sizes[i - 1])
    assert total.value == sum(sizes) == 5 * (4096 * 4096 * 2 + 2 * 1024 * 4096 + 3 * 12288 * 4096 + 2 * 128 + 2 * 4096) + 4096 * 5 * 4096 + 2 * 4096
    ws = L.sf_dflash_workspace_bytes(cfg)
    assert 20e9 < ws < 80e9, ws
    bad = SfDflashConfig(4, 2048, 512, 16, 4096, 5, 12288, 32, 8, 100, 5, 151936, 151669, 41000, 1e-6, 0.0)
    assert L.sf_dflash_workspace_bytes(bad) == 0 and b"multiples of 8" in L.sf_last_error()
    # sliding-window layers (configs/qwen3.6-27b-dflash.json: 4 sliding + 1 full, W = 2048): the layout fields validate like
    # resolve_dflash_attention_layout (modeling/draft/dflash.py:38-68)
    base = (4, 2048, 512, 16, 4096, 5, 12288, 32, 8, 128, 5, 151936, 151669, 41000, 1e-6, 0.0, 0, 0, 0.5)
    assert L.sf_dflash_num_params(SfDflashConfig(*base, 2048, 0b01111)) == n
    assert L.sf_dflash_num_params(SfDflashConfig(*base, 0, 0b00001)) < 0 and b"positive sliding_window" in L.sf_last_error()
    assert L.sf_dflash_num_params(SfDflashConfig(*base, 2048, 0b100000)) < 0 and b"num_layers" in L.sf_last_error()


def test_debug_option_names(L):
    L.sf_debug_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    for name in (b"no_pdl", b"loss_side", b"no_overlap", b"no_swiglu_fusion", b"gemm_group_m", b"gemm_group_m_midk", b"gemm_group_m_wgrad",
                 b"dflash_attn_tc", b"gemm_stages", b"no_teacher_fusion", b"no_loss_stats_fusion", b"gemm_wide", b"gemm_epi_staged",
                 b"no_rope_fusion", b"gemm_epi8", b"dflash_attn_window"):
        assert L.sf_debug_option(name, 0) == 0, name
    assert L.sf_debug_option(b"no_such_switch", 1) != 0
    assert b"unknown option" in L.sf_last_error()
