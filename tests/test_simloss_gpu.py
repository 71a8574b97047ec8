"""GPU parity tests of cosine_scores / make_labels / sent_loss / img_loss: the CUDA path through
the C ABI against (a) the committed golden vectors produced by the reference's own functions and
(b) the CPU oracle in float64 on the same seeded inputs.  Tolerances are north_star's."""
import glob
import math
import os

import numpy as np
import pytest
import torch

import oracle
from util import TOL_BF16, TOL_FP32, lerr, nerr, planted_sent, t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    from xmc_gan_b200 import train_gan
    return train_gan


def _run(T, fn, a, b, labels, b_global, need=(True, True), scale_out=1.0, **kw):
    a = a.clone().cuda().requires_grad_(need[0])
    b = b.clone().cuda().requires_grad_(need[1])
    loss = fn(a, b, labels, b_global, **kw)
    (loss * scale_out).backward()
    return loss.detach().cpu(), a.grad, b.grad


def _oracle(fn, a, b, labels, b_global, smooth, need=(True, True), scale_out=1.0, scale=1.0):
    a = a.double().clone().requires_grad_(need[0])
    b = b.double().clone().requires_grad_(need[1])
    if scale != 1.0:
        loss = oracle.infonce_tail(scale * oracle.cosine_scores(a, b), labels.double(),
                                   oracle.num_pos_of(labels, b_global, smooth))
    else:
        loss = fn(a, b, labels, b_global, smooth)
    (loss * scale_out).backward()
    return loss.detach(), a.grad, b.grad


def test_golden_reference_vectors(T, golden_dir):
    """Outputs of the REFERENCE's own code (tests/golden/ref_*.npz) vs the CUDA path."""
    files = sorted(p for p in glob.glob(os.path.join(golden_dir, "ref_*.npz")) if "ref_magp_" not in p and "ref_attn_" not in p)
    assert files
    for path in files:
        g = np.load(path)
        kind, b_global, smooth = str(g["kind"]), bool(g["b_global"]), float(g["smooth_global"])
        need = tuple(bool(v) for v in g["need"])
        T.cfg.TRAIN.SMOOTH.GLOBAL = smooth
        labels = T.make_labels(g["a"].shape[0], t(g["sent"]).cuda(), b_global)
        assert torch.equal(labels.cpu(), t(g["labels"])), path
        fn = T.sent_loss if kind == "sent" else T.img_loss
        loss, da, db = _run(T, fn, t(g["a"]), t(g["b"]), labels, b_global, need)
        assert lerr(loss, g["loss64"]) <= TOL_FP32, (path, float(loss), float(g["loss64"]))
        if need[0]:
            assert nerr(da, t(g["da64"])) <= TOL_FP32, path
        else:
            assert da is None
        if need[1]:
            assert nerr(db, t(g["db64"])) <= TOL_FP32, path
        else:
            assert db is None
        sc = T.cosine_scores(t(g["a"]).cuda(), t(g["b"]).cuda())
        assert torch.allclose(sc.cpu(), t(g["scores"]), atol=2e-6), path
    T.cfg.TRAIN.SMOOTH.GLOBAL = 0.5


@pytest.mark.parametrize("B,D", [(32, 256), (256, 256), (256, 512), (88, 256), (61, 128), (7, 64), (40, 768)])
@pytest.mark.parametrize("b_global,smooth", [(False, 0.5), (True, 0.5), (True, 0.0)])
def test_sent_loss_vs_oracle(T, B, D, b_global, smooth):
    g = torch.Generator().manual_seed(B * 1000 + D)
    a = torch.randn(B, D, generator=g)
    b = torch.randn(B, D, generator=g) + 0.4 * a
    sent = planted_sent(B, 64, g)
    T.cfg.TRAIN.SMOOTH.GLOBAL = smooth
    labels = T.make_labels(B, sent.cuda(), b_global)
    lab_o = oracle.make_labels(B, sent, b_global, smooth_global=smooth)
    assert torch.equal(labels.cpu(), lab_o)
    loss, da, db = _run(T, T.sent_loss, a, b, labels, b_global, scale_out=2.5)
    lo, dao, dbo = _oracle(oracle.sent_loss, a, b, lab_o, b_global, smooth, scale_out=2.5)
    T.cfg.TRAIN.SMOOTH.GLOBAL = 0.5
    assert lerr(loss, lo) <= TOL_FP32
    assert nerr(da, dao) <= TOL_FP32 and nerr(db, dbo) <= TOL_FP32


def test_dense_identity_equals_tagged_identity(T):
    """An untagged dense eye(B) (what a user might pass) gives the same result as the NULL path."""
    g = torch.Generator().manual_seed(5)
    a, b = torch.randn(64, 256, generator=g), torch.randn(64, 256, generator=g)
    l1, da1, db1 = _run(T, T.sent_loss, a, b, T.make_labels(64, None, False), False)
    l2, da2, db2 = _run(T, T.sent_loss, a, b, torch.eye(64).cuda(), False)
    assert lerr(l1, l2) < 1e-6 and nerr(da1, da2) < 1e-6 and nerr(db1, db2) < 1e-6


def test_img_loss_only_fake_grad(T):
    """img_loss call site: real detached, fake needs grad (train_gan.py:271-278), D=512."""
    g = torch.Generator().manual_seed(11)
    a, b = torch.randn(256, 512, generator=g), torch.randn(256, 512, generator=g)
    labels = T.make_labels(256, None, False)
    loss, da, db = _run(T, T.img_loss, a, b, labels, False, need=(False, True))
    lo, _, dbo = _oracle(oracle.img_loss, a, b, torch.eye(256), False, 0.5, need=(False, True))
    assert da is None and lerr(loss, lo) <= TOL_FP32 and nerr(db, dbo) <= TOL_FP32


def test_temperature_keyword(T):
    g = torch.Generator().manual_seed(12)
    a, b = torch.randn(96, 256, generator=g), torch.randn(96, 256, generator=g)
    b = b + 0.5 * a
    labels = T.make_labels(96, None, False)
    loss, da, db = _run(T, T.sent_loss, a, b, labels, False, tau=0.1)
    lo, dao, dbo = _oracle(None, a, b, torch.eye(96), False, 0.5, scale=10.0)
    assert lerr(loss, lo) <= TOL_FP32 and nerr(da, dao) <= TOL_FP32 and nerr(db, dbo) <= TOL_FP32


def test_bf16_inputs(T):
    """bf16 storage, fp32 arithmetic: oracle is fed the bf16-rounded values."""
    g = torch.Generator().manual_seed(13)
    a = torch.randn(256, 256, generator=g).bfloat16()
    b = (torch.randn(256, 256, generator=g) + 0.4 * a.float()).bfloat16()
    labels = T.make_labels(256, None, False)
    loss, da, db = _run(T, T.sent_loss, a, b, labels, False)
    lo, dao, dbo = _oracle(oracle.sent_loss, a.float(), b.float(), torch.eye(256), False, 0.5)
    assert da.dtype == torch.bfloat16
    assert lerr(loss, lo) <= TOL_BF16 and nerr(da, dao) <= TOL_BF16 and nerr(db, dbo) <= TOL_BF16


def test_known_answers(T):
    B = 16
    x = torch.randn(1, 128).repeat(B, 1).cuda()
    lab = T.make_labels(B, None, False)
    assert abs(float(T.sent_loss(x, x, lab, False)) - 2 * math.log(B)) < 1e-5
    q = torch.linalg.qr(torch.randn(128, 128, dtype=torch.float64))[0][:B].float().cuda()
    assert abs(float(T.sent_loss(q, q, lab, False)) - 2 * math.log(1 + (B - 1) / math.e)) < 1e-5


def test_edge_cases(T):
    """B=1, a zero row (norm clamped to eps as F.normalize does), rectangular cosine_scores."""
    a = torch.randn(1, 256).cuda().requires_grad_()
    loss = T.sent_loss(a, a.detach().clone().requires_grad_(), T.make_labels(1, None, False), False)
    loss.backward()
    assert abs(float(loss)) < 1e-6 and torch.isfinite(a.grad).all()
    g = torch.Generator().manual_seed(14)
    x, y = torch.randn(20, 256, generator=g), torch.randn(33, 256, generator=g)
    x[3] = 0
    sc = T.cosine_scores(x.cuda(), y.cuda())
    assert sc.shape == (20, 33)
    assert torch.allclose(sc.cpu(), oracle.cosine_scores(x, y), atol=2e-6)
    xa = x[:20].clone().cuda().requires_grad_()
    yb = y[:20].clone().cuda().requires_grad_()
    T.sent_loss(xa, yb, T.make_labels(20, None, False), False).backward()
    assert torch.isfinite(xa.grad).all() and torch.isfinite(yb.grad).all()
    xo, yo = x[:20].double().requires_grad_(), y[:20].double().requires_grad_()
    oracle.sent_loss(xo, yo, torch.eye(20), False).backward()
    assert nerr(yb.grad, yo.grad) <= TOL_FP32
    keep = torch.arange(20) != 3
    assert nerr(xa.grad[keep], xo.grad[keep]) <= TOL_FP32


def test_full_size_properties(T):
    """BASELINE config sizes (256 and 2048 columns): symmetry and invariances, no oracle needed."""
    g = torch.Generator().manual_seed(15)
    a, b = torch.randn(256, 256, generator=g).cuda(), torch.randn(256, 256, generator=g).cuda()
    lab = T.make_labels(256, None, False)
    l_ab = float(T.sent_loss(a, b, lab, False))
    l_ba = float(T.sent_loss(b, a, lab, False))                    # s0 <-> s1 swap
    assert abs(l_ab - l_ba) < 1e-5 * abs(l_ab)
    l_sc = float(T.sent_loss(3.0 * a, 0.25 * b, lab, False))       # cosine is scale invariant
    assert abs(l_ab - l_sc) < 1e-5 * abs(l_ab)
    perm = torch.randperm(256, generator=g).cuda()
    l_pm = float(T.sent_loss(a[perm], b[perm], lab, False))        # joint permutation
    assert abs(l_ab - l_pm) < 1e-5 * abs(l_ab)
    ar = a.clone().requires_grad_()
    T.sent_loss(ar, b, lab, False).backward()
    assert float((ar.grad * ar.detach()).sum(1).abs().max()) < 1e-5   # grad orthogonal to the row (normalise)


# ---- large rectangular problems: the tcgen05 form (simloss_tc.cu) -------------------------------------------------
@pytest.mark.parametrize("Bq,Bk,D", [(256, 2048, 256), (256, 2048, 512), (200, 1320, 256), (256, 1024, 768), (130, 2056, 128)])
@pytest.mark.parametrize("dt,tol", [(torch.float32, TOL_FP32), (torch.bfloat16, TOL_BF16)])
@pytest.mark.parametrize("dense", [False, True])
def test_large_rectangular_tensor_core_path_vs_oracle(T, Bq, Bk, D, dt, tol, dense):
    """Rank-shaped problem (local rows x gathered columns, identity labels with a diagonal offset or dense soft
    labels): split-bf16 tcgen05 tiles must stay inside the fp32 tolerance; ragged tiles, D up to 768."""
    from xmc_gan_b200 import _lib
    from xmc_gan_b200.ops import default_ops
    ops = default_ops()
    assert _lib.lib().xmc_simloss_workspace_bytes(Bq, Bk, D) > 0, "shape must take the tensor-core form"
    g = torch.Generator().manual_seed(Bq + Bk + D)
    a = torch.randn(Bq, D, generator=g).to(dt)
    b = (torch.randn(Bk, D, generator=g) + 0.4 * a.float()[torch.randint(0, Bq, (Bk,), generator=g)]).to(dt)
    diag = 64 if Bk >= Bq + 64 else 0
    lab_o = torch.zeros(Bq, Bk)
    lab_o[torch.arange(Bq), torch.arange(Bq) + diag] = 1.0
    if dense:
        extra = torch.rand(Bq, Bk, generator=g) < 0.002
        lab_o = (lab_o + 0.3 * extra.float()).clamp(max=1.0)
    labels = lab_o.cuda() if dense else None
    scale = 2.0
    go = torch.tensor(1.7, device="cuda")
    ac, bc = a.cuda(), b.cuda()
    scores, inv_a, inv_b, row_stats, col_stats = ops.simloss_forward(ac, bc, labels, diag, scale)
    loss3 = ops.infonce_loss(row_stats, col_stats, None, None, 1.0, Bq, Bk, 0, Bk)
    da, db = ops.simloss_backward(ac, bc, scores, inv_a, inv_b, labels, diag, scale, row_stats, col_stats, None, None, 1.0,
                                  Bq, Bk, go, True, True)
    ar, br = a.double().requires_grad_(), b.double().requires_grad_()
    so = oracle.cosine_scores(ar, br)
    lo = oracle.infonce_tail(scale * so, lab_o.double(), 1)
    (1.7 * lo).backward()
    assert nerr(scores, so) <= 1e-5                                         # split-bf16 product: ~2^-17 per operand
    assert lerr(loss3[0], lo.detach()) <= tol
    assert nerr(da, ar.grad) <= tol and nerr(db, br.grad) <= tol
    assert da.dtype == dt and db.dtype == dt
    # one-sided gradients (img_loss: only db; D step: only da) and the CUDA-core form on the same problem
    _, db1 = ops.simloss_backward(ac, bc, scores, inv_a, inv_b, labels, diag, scale, row_stats, col_stats, None, None, 1.0,
                                  Bq, Bk, go, False, True)
    assert nerr(db1, br.grad) <= tol
    ops.use_sim_tc = False
    try:
        da0, db0 = ops.simloss_backward(ac, bc, scores, inv_a, inv_b, labels, diag, scale, row_stats, col_stats, None, None,
                                        1.0, Bq, Bk, go, True, True)
    finally:
        ops.use_sim_tc = True
    assert nerr(da, da0) <= tol and nerr(db, db0) <= tol
