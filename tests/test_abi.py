"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol that
include/xmc_loss.h declares; the Python binding lists exactly those symbols; the product path
refuses to run without CUDA (no fallback).  No compute calls here — there is no GPU."""
import ctypes
import os

import pytest
import torch

from xmc_gan_b200 import _lib


@pytest.fixture(scope="module")
def built():
    return _lib.build()


def test_header_and_binding_agree():
    assert _lib.header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(built):
    h = ctypes.CDLL(built)
    for name in _lib.header_symbols():
        assert hasattr(h, name), f"{name} declared in include/xmc_loss.h but not exported"


def test_version_and_error_string(built):
    L = _lib.lib()
    assert L.xmc_version() == 6
    assert isinstance(L.xmc_last_error(), bytes)


def test_argument_validation_without_gpu(built):
    """Bad arguments are rejected before any CUDA call, with a message."""
    L = _lib.lib()
    rc = L.xmc_cosine_scores(None, None, 4, 4, 8, 0, None, None, None, None)
    assert rc == 1 and b"null" in L.xmc_last_error()
    rc = L.xmc_wordregion_forward(0, 16, 16, None, 8, 2, 5, 7, 64, 5.0, 16, 16, 16, None, None, None, 0, None)
    assert rc == 1 and b"Rpad" in L.xmc_last_error()
    rc = L.xmc_wordregion_forward(0, 16, 16, None, 8, 2, 5, 16, 96, 5.0, 16, 16, 16, None, None, None, 0, None)
    assert rc == 2 and b"unsupported" in L.xmc_last_error()
    rc = L.xmc_cosine_scores(8, 16, 4, 4, 8, 0, 16, None, None, None)
    assert rc == 3
    # MA-GP reduction: empty problem, missing outputs, power below 2
    assert L.xmc_gradnorm_penalty_forward(None, 0, None, 0, 4, 0, 6.0, 2.0, 1, 16, 16, 16, None) == 1
    assert L.xmc_gradnorm_penalty_forward(16, 8, None, 0, 4, 0, 6.0, 2.0, 1, None, 16, 16, None) == 1
    assert L.xmc_gradnorm_penalty_forward(16, 8, None, 0, 4, 0, 1.0, 2.0, 1, 16, 16, 16, None) == 1
    assert b"power" in L.xmc_last_error()
    assert L.xmc_gradnorm_penalty_backward(16, 8, 16, 8, 4, 0, 6.0, 2.0, 1, None, 16, 16, 16, None) == 1
    # region head / pooled embedding: null pointers, widths other than 256, Rpad not a multiple of 16
    assert L.xmc_region_head_forward(None, 0, 16, 0, None, 2, 64, 16, 16, 256, 16, 16, None) == 1
    assert L.xmc_region_head_forward(16, 0, 16, 0, None, 2, 64, 16, 16, 128, 16, 16, None) == 2 and b"D=128" in L.xmc_last_error()
    assert L.xmc_region_head_forward(16, 0, 16, 0, None, 2, 64, 16, 20, 256, 16, 16, None) == 1
    assert L.xmc_region_head_forward(16, 7, 16, 0, None, 2, 64, 16, 16, 256, 16, 16, None) == 2
    assert L.xmc_region_head_backward_input(16, 0, None, 0, 2, 64, 16, 256, 16, 0, None) == 1
    assert L.xmc_region_head_backward_weight(16, 0, 16, 0, 0, 64, 16, 256, 16, None, None) == 1
    assert L.xmc_region_head_backward_weight(8, 0, 16, 0, 2, 64, 16, 256, 16, None, None) == 3
    assert L.xmc_avgpool_rows(None, 0, 2, 8, 16, 16, 0, None) == 1
    assert L.xmc_avgpool_rows(16, 0, 2, 8, 0, 16, 0, None) == 1
    assert L.xmc_avgpool_rows_backward(16, 5, 2, 8, 16, 16, 0, None) == 2
    # merge of per-rank column statistics
    assert L.xmc_infonce_combine_stats(None, 2, 8, 16, None) == 1
    assert L.xmc_infonce_combine_stats(16, 0, 8, 16, None) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    from xmc_gan_b200 import train_gan as T
    a = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.sent_loss(a, a, torch.eye(4), False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.cosine_scores(a, a)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="missing"):
        _lib.lib()
