"""CPU: the C-ABI library loads and exports every symbol include/howl_b200.h declares; host integer helpers."""
import ctypes as C
import os

import numpy as np
import pytest

from howl_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def test_every_header_symbol_is_exported_and_bound(lib):
    names = _lib.header_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} declared in the header but not bound in howl_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    # the drop-in boundary (howl_b200.h) carries no tuning aids / test hooks: those live in howl_b200_debug.h
    product = _lib.header_symbols(debug=False)
    assert not [n for n in product if "debug" in n or "selftest" in n]


def test_abi_version(lib):
    assert lib.howl_b200_abi_version() == 1


def test_num_frames_and_compute_lengths_bit_exact(lib, golden):
    g = golden("frontend")
    assert lib.howl_b200_num_frames(8000, 200) == 41
    assert lib.howl_b200_num_frames(16000, 200) == 81
    assert lib.howl_b200_num_frames(0, 200) == 1
    lens = np.ascontiguousarray(g["lengths_in"], dtype=np.int64)
    out = np.zeros_like(lens)
    rc = lib.howl_b200_compute_lengths(lens.ctypes.data_as(C.c_void_p), lens.size, 512, 200, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.array_equal(out, g["lengths_out"])


def test_param_count_and_workspace(lib):
    assert lib.howl_b200_res8_param_count(4) == 109939
    assert lib.howl_b200_res8_param_count(30) == 111135
    assert lib.howl_b200_res8_workspace_bytes(64, 41, 40, 4, 1) > 0
    assert lib.howl_b200_res8_workspace_bytes(64, 41, 80, 4, 1) == -1


def test_create_without_gpu_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.FrontendCfg(16000, 512, 200, 40)
    h = C.c_void_p()
    rc = lib.howl_b200_create(0, C.byref(cfg), C.byref(h))
    assert rc < 0
    assert b"no CPU fallback" in lib.howl_b200_last_error(None)
    import howl_b200

    with pytest.raises(howl_b200.HowlB200Error):
        howl_b200.Context("cuda:0")
