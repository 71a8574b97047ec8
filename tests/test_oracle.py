"""CPU tests that pin the oracle: golden vectors generated from the REFERENCE's own functions
(tests/golden/ref_*.npz, made by tests/golden/make_golden.py), the reference itself when
/root/reference is present (build container only), and the known-answer tests of SURVEY §4."""
import glob
import math
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import load_reference as LR


def _t(x, dtype=None):
    t = torch.from_numpy(np.asarray(x))
    return t.to(dtype) if dtype is not None else t


def _ref_cases(golden_dir):
    return sorted(p for p in glob.glob(os.path.join(golden_dir, "ref_*.npz")) if "ref_magp_" not in p and "ref_attn_" not in p)


def test_golden_files_present(golden_dir):
    assert len(_ref_cases(golden_dir)) >= 6
    assert len(glob.glob(os.path.join(golden_dir, "word_*.npz"))) >= 4


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 1e-12)])
def test_oracle_matches_reference_golden(golden_dir, dtype, tol):
    """Oracle restatement == reference outputs (loss, both gradients, labels, cosine matrix)."""
    for path in _ref_cases(golden_dir):
        g = np.load(path)
        kind, b_global, smooth = str(g["kind"]), bool(g["b_global"]), float(g["smooth_global"])
        need = [bool(v) for v in g["need"]]
        a = _t(g["a"], dtype).requires_grad_(need[0])
        b = _t(g["b"], dtype).requires_grad_(need[1])
        labels = oracle.make_labels(a.shape[0], _t(g["sent"]), b_global, smooth_global=smooth)
        assert torch.equal(labels, _t(g["labels"])), path           # bit-exact label matrix
        fn = oracle.sent_loss if kind == "sent" else oracle.img_loss
        loss = fn(a, b, labels, b_global, smooth)
        loss.backward()
        sfx = "32" if dtype == torch.float32 else "64"
        assert abs(float(loss) - float(g["loss" + sfx])) <= tol * max(1.0, abs(float(g["loss" + sfx]))), path
        for t, key, n in ((a, "da", need[0]), (b, "db", need[1])):
            if n:
                ref = _t(g[key + sfx])
                err = (t.grad - ref).norm() / ref.norm()
                assert err <= (5e-6 if dtype == torch.float32 else 1e-11), (path, key, float(err))
        if dtype == torch.float32:
            sc = oracle.cosine_scores(_t(g["a"]), _t(g["b"]))
            assert torch.allclose(sc, _t(g["scores"]), atol=1e-6), path


@pytest.mark.skipif(not LR.reference_available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("b_global,smooth", [(False, 0.5), (True, 0.5), (True, 0.0)])
def test_oracle_matches_reference_live(b_global, smooth):
    """Same seeded inputs through the AST-loaded reference source and through the oracle."""
    ref = LR.load_reference_losses(smooth)
    g = torch.Generator().manual_seed(1234)
    B, D = 48, 128
    a = torch.randn(B, D, generator=g, dtype=torch.float64)
    b = torch.randn(B, D, generator=g, dtype=torch.float64) + 0.3 * a
    sent = torch.randn(B, 32, generator=g)
    sent[5] = sent[9] + 0.05 * torch.randn(32, generator=g)
    sent[17] = sent[9] + 0.05 * torch.randn(32, generator=g)
    with LR.cuda_is_identity():
        lab_ref = ref.make_labels(B, sent, b_global)
    lab = oracle.make_labels(B, sent, b_global, smooth_global=smooth)
    assert torch.equal(lab, lab_ref)
    for name in ("sent_loss", "img_loss"):
        a1, b1 = a.clone().requires_grad_(), b.clone().requires_grad_()
        a2, b2 = a.clone().requires_grad_(), b.clone().requires_grad_()
        l_ref = getattr(ref, name)(a1, b1, lab_ref, b_global)
        l_orc = getattr(oracle, name)(a2, b2, lab, b_global, smooth)
        l_ref.backward(); l_orc.backward()
        assert abs(float(l_ref) - float(l_orc)) < 1e-12
        assert torch.allclose(a1.grad, a2.grad, atol=1e-13) and torch.allclose(b1.grad, b2.grad, atol=1e-13)


def test_known_answers():
    """identical rows -> 2 ln B; orthonormal rows, identity labels -> 2 ln(1+(B-1)/e)."""
    B = 16
    x = torch.randn(1, 32, dtype=torch.float64).repeat(B, 1)
    eye = torch.eye(B)
    assert abs(float(oracle.sent_loss(x, x, eye, False)) - 2 * math.log(B)) < 1e-12
    q = torch.linalg.qr(torch.randn(64, 64, dtype=torch.float64))[0][:B]
    assert abs(float(oracle.sent_loss(q, q, eye, False)) - 2 * math.log(1 + (B - 1) / math.e)) < 1e-12
    a, b = torch.randn(B, 24, dtype=torch.float64), torch.randn(B, 24, dtype=torch.float64)
    S = oracle.cosine_scores(a, b)
    tgt = torch.arange(B)
    ce = torch.nn.functional.cross_entropy
    assert abs(float(oracle.sent_loss(a, b, eye, False)) - float(ce(S, tgt) + ce(S.t(), tgt))) < 1e-12


def test_closed_form_gradient_of_tail():
    """The closed-form dS the CUDA kernels use equals autograd of the tail."""
    g = torch.Generator().manual_seed(7)
    Bq = 12
    S = (torch.rand(Bq, Bq, generator=g, dtype=torch.float64) * 2 - 1).requires_grad_()
    L = (torch.rand(Bq, Bq, generator=g) > 0.7).double() * 0.5 + torch.eye(Bq, dtype=torch.float64)
    L = L.clamp(max=1)
    n = (L > 0).sum(1)
    loss = oracle.infonce_tail(S, L, n)
    loss.backward()
    Sd = S.detach()
    pc, pr = torch.softmax(Sd, 0), torch.softmax(Sd, 1)
    nr, nc = n.double().view(-1, 1), n.double().view(1, -1)       # column j divided by n[j] (row count of row j)
    dS = ((pc * L.sum(0, keepdim=True) - L) / nc + (pr * L.sum(1, keepdim=True) - L) / nr) / Bq
    assert torch.allclose(S.grad, dS, atol=1e-14)


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)


@pytest.mark.parametrize("dtype,tag,tol", [(torch.float32, "32", 2e-6), (torch.float64, "64", 1e-13)])
def test_attention_stage_matches_reference_golden(golden_dir, dtype, tag, tol):
    """oracle.attend (cosines -> softmax over the regions -> contexts) against the reference's own attention block
    (concept_gan.py:532-555) on every (image, caption) pair: rho1 = 1, unit values."""
    paths = sorted(glob.glob(os.path.join(golden_dir, "ref_attn_*.npz")))
    assert len(paths) >= 2
    for path in paths:
        g = np.load(path)
        en = _unit(_t(g["words"], dtype).transpose(1, 2))
        vn = _unit(_t(g["regions"], dtype).transpose(1, 2))
        _, a, ctx = oracle.attend(en, vn, vn, 1.0)
        ref = _t(g["ctx" + tag], dtype)
        assert float((ctx - ref).abs().max()) <= tol, path
        assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=10 * tol)
        if tag == "64":        # masked keys (the -inf convention) == attending over the unmasked prefix only
            m = int(g["masked_from"])
            _, _, ctx_m = oracle.attend(en, vn[:, :m], vn[:, :m], 1.0)
            assert float((ctx_m - _t(g["ctx64_masked"], dtype)).abs().max()) <= tol, path


@pytest.mark.skipif(not LR.reference_available(), reason="needs /root/reference (build container only)")
def test_attention_stage_matches_reference_live():
    ref = LR.load_reference_attention()
    g = torch.Generator().manual_seed(9)
    Bi, Bc, D, T, R = 2, 3, 32, 5, 17
    words = torch.randn(Bc, D, T, generator=g, dtype=torch.float64)
    regions = torch.randn(Bi, D, R, generator=g, dtype=torch.float64)
    q = words.unsqueeze(0).expand(Bi, Bc, D, T).reshape(Bi * Bc, D, T).clone()
    k = regions.unsqueeze(1).expand(Bi, Bc, D, R).reshape(Bi * Bc, D, R).clone()
    want = ref(q, k, torch.zeros(Bi * Bc, R, dtype=torch.bool)).reshape(Bi, Bc, T, D)
    en, vn = _unit(words.transpose(1, 2)), _unit(regions.transpose(1, 2))
    _, _, ctx = oracle.attend(en, vn, vn, 1.0)
    assert float((ctx - want).abs().max()) <= 1e-13


def test_word_oracle_regression(golden_dir):
    """word_*.npz are outputs of this repo's restatement (parity UNPINNED by the reference)."""
    for path in sorted(glob.glob(os.path.join(golden_dir, "word_*.npz"))):
        g = np.load(path)
        rho = [float(v) for v in g["rho"]]
        w = _t(g["words"], torch.float64).requires_grad_()
        r = _t(g["regions"], torch.float64).requires_grad_()
        mask, labels = _t(g["mask"]), _t(g["labels"])
        nv, bg, sm = bool(g["normalize_values"]), bool(g["b_global"]), float(g["smooth_global"])
        S = oracle.word_scores(r, w, mask, rho[0], rho[1], nv)
        loss = oracle.word_loss(r, w, mask, labels, bg, sm, rho[0], rho[1], rho[2], nv)
        loss.backward()
        assert torch.allclose(S.detach(), _t(g["scores"]), atol=1e-6), path
        assert abs(float(loss) - float(g["loss"])) < 1e-6, path
        for t, key in ((w, "dwords"), (r, "dregions")):
            ref = _t(g[key], torch.float64)
            assert (t.grad - ref).norm() / ref.norm() < 1e-5, (path, key)
        assert torch.isfinite(w.grad).all() and torch.isfinite(r.grad).all()


def test_word_oracle_properties():
    """Size-independent properties of the spec: block independence, padding invariance, bounds."""
    g = torch.Generator().manual_seed(3)
    B, D, T, R = 5, 32, 6, 11
    w = torch.randn(B, D, T, generator=g, dtype=torch.float64)
    r = torch.randn(B, D, R, generator=g, dtype=torch.float64)
    lens = torch.tensor([6, 3, 1, 4, 2])
    mask = torch.arange(T).unsqueeze(0) >= lens.unsqueeze(1)
    S = oracle.word_scores(r, w, mask, 5.0, 5.0)
    assert torch.allclose(S, oracle.word_scores(r, w, mask, 5.0, 5.0, img_block=2), atol=1e-14)
    # values of padded words do not matter
    w2 = w.clone(); w2[mask.unsqueeze(1).expand_as(w2)] = 123.0
    assert torch.allclose(S, oracle.word_scores(r, w2, mask, 5.0, 5.0), atol=1e-12)
    # |rel| <= 1  =>  -1 <= S <= 1 + ln(len)/rho2
    assert (S >= -1 - 1e-9).all() and (S <= 1 + torch.log(lens.double()).view(1, -1) / 5.0 + 1e-9).all()
    # scaling words or regions by positive constants leaves normalised-value scores unchanged
    Sn = oracle.word_scores(r, w, mask, 5.0, 5.0, True)
    assert torch.allclose(Sn, oracle.word_scores(3.0 * r, 0.5 * w, mask, 5.0, 5.0, True), atol=1e-12)
    # a fully padded caption scores 0 and has zero, finite gradient
    mask[2] = True
    wq = w.clone().requires_grad_()
    S2 = oracle.word_scores(r, wq, mask, 5.0, 5.0)
    assert (S2[:, 2] == 0).all()
    S2.sum().backward()
    assert torch.isfinite(wq.grad).all() and (wq.grad[2] == 0).all()


# ---- MA-GP reduction (train_gan.py:244-249) ---------------------------------------------------------
def _magp_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "ref_magp_*.npz")))


@pytest.mark.parametrize("dtype,tag,tol", [(torch.float32, "32", 2e-6), (torch.float64, "64", 1e-12)])
def test_magp_oracle_matches_reference_golden(golden_dir, dtype, tag, tol):
    cases = _magp_cases(golden_dir)
    assert len(cases) >= 2
    for path in cases:
        z = np.load(path)
        a = torch.from_numpy(z["g0"]).to(dtype).requires_grad_()
        b = torch.from_numpy(z["g1"]).to(dtype).requires_grad_()
        loss = oracle.magp_penalty(a, b)
        loss.backward()
        assert abs(float(loss) - float(z["loss" + tag])) <= tol * abs(float(z["loss" + tag])), path
        for got, ref in ((a.grad, z["d0_" + tag]), (b.grad, z["d1_" + tag])):
            ref = torch.from_numpy(ref)
            assert float((got - ref).norm() / ref.norm()) <= 10 * tol, path


@pytest.mark.skipif(not LR.reference_available(), reason="/root/reference only exists in the build container")
def test_magp_oracle_matches_reference_live():
    ref = LR.load_reference_magp()
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(7, 3, 5, 6, generator=g, dtype=torch.float64) * 0.1).requires_grad_()
    b = (torch.randn(7, 12, generator=g, dtype=torch.float64) * 0.1).requires_grad_()
    l_ref = ref((a, b))
    g_ref = torch.autograd.grad(l_ref, (a, b))
    l_or = oracle.magp_penalty(a, b)
    g_or = torch.autograd.grad(l_or, (a, b))
    assert abs(float(l_ref) - float(l_or)) <= 1e-13 * abs(float(l_ref))
    for x, y in zip(g_ref, g_or):
        assert torch.allclose(x, y, rtol=1e-12, atol=0)
