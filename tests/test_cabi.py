"""The C-ABI shared library: loads without a GPU, exports every symbol the header declares,
and refuses (loudly) to create a model when there is no sm_100 device."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from vit_flax_b200 import _lib

HEADER = Path(__file__).resolve().parents[1] / "include" / "vitb200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(vitb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib_built):
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.SIGNATURES) == syms


def test_every_declared_symbol_is_exported(lib_built):
    raw = C.CDLL(str(_lib.LIB_PATH))
    for s in declared_symbols():
        assert hasattr(raw, s), f"{s} declared in include/vitb200.h but not exported"
    assert lib_built.vitb200_abi_version() == _lib.ABI_VERSION == 4


def test_config_struct_layout():
    assert C.sizeof(_lib.Config) == 18 * 4


def test_create_without_gpu_is_an_error_not_a_fallback(lib_built):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.Config(image_h=32, image_w=32, patch_h=8, patch_w=8, channels=3, num_classes=8,
                      dim=64, depth=1, heads=1, mlp_dim=64, pool=0, precision=0, max_batch=1)
    h = C.c_void_p()
    rc = lib_built.vitb200_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -5 and not h.value
    assert b"no CUDA device" in lib_built.vitb200_last_error()
    with pytest.raises(_lib.VitB200Error):
        _lib.check(rc)


def test_create_rejects_bad_config_before_touching_the_device(lib_built):
    cfg = _lib.Config(image_h=30, image_w=32, patch_h=8, patch_w=8, channels=3, num_classes=8,
                      dim=64, depth=1, heads=1, mlp_dim=64, pool=0, precision=0, max_batch=1)
    h = C.c_void_p()
    assert lib_built.vitb200_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"divisible" in lib_built.vitb200_last_error()
