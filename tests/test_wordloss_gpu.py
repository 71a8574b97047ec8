"""GPU parity tests of word_loss (word–region attention contrastive loss) through the C ABI.

The reference does not implement this loss (train_gan.py:220-222, 267-269): the checker is this
repo's CPU restatement (oracle/word_region.py, PARITY UNPINNED) in float64 on the same seeded
inputs, plus the committed word_*.npz fixtures and size-independent properties at full size."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from util import TOL_BF16, TOL_FP32, lerr, nerr, planted_sent, t, word_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    from xmc_gan_b200 import train_gan
    return train_gan


def _gpu(T, regions, words, mask, labels, b_global, **kw):
    r = regions.clone().cuda().requires_grad_()
    w = words.clone().cuda().requires_grad_()
    loss = T.word_loss(r, w, None if mask is None else mask.cuda(), labels, b_global, **kw)
    loss.backward()
    return loss.detach().cpu(), r.grad, w.grad


def _cpu(regions, words, mask, labels, b_global, smooth=0.5, rho=(5.0, 5.0, 10.0), nv=False):
    r = regions.double().clone().requires_grad_()
    w = words.double().clone().requires_grad_()
    loss = oracle.word_loss(r, w, mask, labels, b_global, smooth, rho[0], rho[1], rho[2], nv)
    loss.backward()
    return loss.detach(), r.grad, w.grad


def _check(res, ref, tol, tag=""):
    (l, dr, dw), (lo, dro, dwo) = res, ref
    assert torch.isfinite(dr).all() and torch.isfinite(dw).all(), tag
    assert lerr(l, lo) <= tol, (tag, float(l), float(lo))
    assert nerr(dr, dro) <= tol, (tag, "d regions", nerr(dr, dro))
    assert nerr(dw, dwo) <= tol, (tag, "d words", nerr(dw, dwo))


def test_golden_fixtures_fp32(T, golden_dir):
    for path in sorted(glob.glob(os.path.join(golden_dir, "word_*.npz"))):
        g = np.load(path)
        rho = [float(v) for v in g["rho"]]
        nv, bg, sm = bool(g["normalize_values"]), bool(g["b_global"]), float(g["smooth_global"])
        T.cfg.TRAIN.SMOOTH.GLOBAL = sm
        labels = t(g["labels"]).cuda()
        res = _gpu(T, t(g["regions"]), t(g["words"]), t(g["mask"]), labels, bg,
                   rho1=rho[0], rho2=rho[1], rho3=rho[2], normalize_values=nv, precision="fp32")
        T.cfg.TRAIN.SMOOTH.GLOBAL = 0.5
        _check(res, (t(g["loss"]), t(g["dregions"]), t(g["dwords"])), TOL_FP32, path)


@pytest.mark.parametrize("B,D,T_,R", [(32, 256, 18, 289), (8, 256, 20, 256), (5, 64, 3, 17), (12, 128, 32, 64),
                                      (3, 256, 12, 100), (16, 256, 18, 64)])
@pytest.mark.parametrize("nv", [False, True])
def test_fp32_vs_oracle(T, B, D, T_, R, nv):
    """CUB-shaped config 1 (B=32, T=18, R=17x17, D=256) and ragged shapes, fp32 path, rel 1e-4."""
    words, regions, mask = word_inputs(B, D, T_, R, seed=B * 100 + R)
    labels = T.make_labels(B, None, False)
    res = _gpu(T, regions, words, mask, labels, False, normalize_values=nv, precision="fp32")
    ref = _cpu(regions, words, mask, torch.eye(B), False, nv=nv)
    _check(res, ref, TOL_FP32, f"B{B} D{D} T{T_} R{R} nv{nv}")


def test_fp32_soft_labels_and_rhos(T):
    B, D, T_, R = 24, 128, 9, 49
    words, regions, mask = word_inputs(B, D, T_, R, seed=77)
    g = torch.Generator().manual_seed(78)
    sent = planted_sent(B, 64, g)
    for smooth in (0.5, 0.0):
        T.cfg.TRAIN.SMOOTH.GLOBAL = smooth
        labels = T.make_labels(B, sent.cuda(), True)
        lab_o = oracle.make_labels(B, sent, True, smooth_global=smooth)
        res = _gpu(T, regions, words, mask, labels, True, rho1=4.0, rho2=6.0, rho3=8.0, precision="fp32")
        ref = _cpu(regions, words, mask, lab_o, True, smooth, rho=(4.0, 6.0, 8.0))
        _check(res, ref, TOL_FP32, f"smooth {smooth}")
    T.cfg.TRAIN.SMOOTH.GLOBAL = 0.5


def test_fp32_4d_regions_no_mask_and_padded_caption(T):
    B, D, T_, H = 6, 64, 8, 5
    words, regions, mask = word_inputs(B, D, T_, H * H, seed=5)
    reg4 = regions.view(B, D, H, H)
    labels = T.make_labels(B, None, False)
    r = reg4.clone().cuda().requires_grad_()
    w = words.clone().cuda().requires_grad_()
    loss = T.word_loss(r, w, None, labels, False, precision="fp32")
    loss.backward()
    assert r.grad.shape == reg4.shape
    ref = _cpu(regions, words, None, torch.eye(B), False)
    _check((loss.detach().cpu(), r.grad.flatten(2), w.grad), ref, TOL_FP32, "no mask")
    mask[2] = True                                            # fully padded caption: score 0, zero grad
    res = _gpu(T, regions, words, mask, labels, False, precision="fp32")
    ref = _cpu(regions, words, mask, torch.eye(B), False)
    _check(res, ref, TOL_FP32, "padded caption")
    assert float(res[2][2].abs().max()) == 0.0


def test_fp32_directional_derivative_full_rows(T):
    """Size-independent check at a larger size: <grad, d> equals a central finite difference."""
    B, D, T_, R = 64, 256, 18, 289
    words, regions, mask = word_inputs(B, D, T_, R, seed=9)
    labels = T.make_labels(B, None, False)
    l, dr, dw = _gpu(T, regions, words, mask, labels, False, precision="fp32")
    g = torch.Generator().manual_seed(10)
    d_r, d_w = torch.randn(regions.shape, generator=g), torch.randn(words.shape, generator=g)
    eps = 2e-2
    with torch.no_grad():
        lp = T.word_loss((regions + eps * d_r).cuda(), (words + eps * d_w).cuda(), mask.cuda(), labels, False, precision="fp32")
        lm = T.word_loss((regions - eps * d_r).cuda(), (words - eps * d_w).cuda(), mask.cuda(), labels, False, precision="fp32")
    fd = (float(lp) - float(lm)) / (2 * eps)
    an = float((dr.cpu() * d_r).sum() + (dw.cpu() * d_w).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1e-3), (fd, an)
