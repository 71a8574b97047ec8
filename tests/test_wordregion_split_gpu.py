"""GPU tests of the fp32-TOLERANCE tensor-core word-region path (wordregion_split.cu, XMC_PATH_FP32_TCGEN05): every fp32
operand carried as a hi + lo bf16 pair, three tcgen05 MMAs per product, fp32 accumulation.  Against the fp32 CUDA-core
kernels on the same operands (tight: the two differ only in the ~2^-17 operand split) and against the fp64 CPU oracle
at north_star's fp32 tolerance, 1e-4."""
import pytest
import torch

import oracle
from util import TOL_FP32, lerr, nerr, word_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from xmc_gan_b200.ops import default_ops
    return default_ops()


def _operands(ops, B, D, T_, R, seed):
    words, regions, mask = word_inputs(B, D, T_, R, seed)
    Rpad = (R + 15) // 16 * 16
    qn, _ = ops.normalize_transpose(words.cuda(), T_, torch.float32)
    kn, rnorm = ops.normalize_transpose(regions.cuda(), Rpad, torch.float32)
    return qn.view(B * T_, D), kn, rnorm


@pytest.mark.parametrize("B,T_,R", [(8, 18, 289), (5, 7, 40), (3, 12, 17), (20, 20, 256), (33, 18, 100)])
@pytest.mark.parametrize("raw_values", [True, False])
def test_kernels_vs_fp32_cuda_core_kernels(ops, B, T_, R, raw_values):
    from xmc_gan_b200 import _lib
    D = 256
    qn, kn, rnorm = _operands(ops, B, D, T_, R, seed=B + R)
    rn = rnorm if raw_values else None
    l1, c1, r1, chat = ops.wordregion_forward(_lib.PATH_FP32_TCGEN05, qn, kn, rn, R, 5.0, save_context=True)
    torch.cuda.synchronize()
    assert int(ops.last_workspace[:4].view(torch.int32)[0]) == 0
    l0, c0, r0, _ = ops.wordregion_forward(_lib.PATH_FP32_SIMT, qn, kn, rn, R, 5.0)
    assert nerr(l1, l0) < 2e-5 and nerr(c1, c0) < 2e-5 and nerr(r1, r0) < 2e-5, (nerr(l1, l0), nerr(c1, c0), nerr(r1, r0))
    g = torch.Generator().manual_seed(1)
    grel = (torch.randn(l1.shape, generator=g) * 0.1).cuda()
    dq1, dk1, dr1 = ops.wordregion_backward(_lib.PATH_FP32_TCGEN05, qn, kn, rn, R, 5.0, l1, c1, r1, grel, chat)
    torch.cuda.synchronize()
    assert int(ops.last_workspace[:4].view(torch.int32)[0]) == 0
    dq0, dk0, dr0 = ops.wordregion_backward(_lib.PATH_FP32_SIMT, qn, kn, rn, R, 5.0, l0, c0, r0, grel)
    assert nerr(dq1, dq0) < 5e-5 and nerr(dk1, dk0) < 5e-5, (nerr(dq1, dq0), nerr(dk1, dk0))
    if raw_values:
        assert nerr(dr1, dr0) < 5e-5, nerr(dr1, dr0)


@pytest.mark.parametrize("B,T_,R", [(16, 18, 289), (7, 5, 33), (40, 12, 256)])
@pytest.mark.parametrize("b_global,smooth", [(False, 0.5), (True, 0.5), (True, 0.0)])
def test_word_loss_fp32_on_tensor_cores_vs_oracle(B, T_, R, b_global, smooth):
    from util import planted_sent
    from xmc_gan_b200 import _lib, losses
    from xmc_gan_b200 import train_gan as T
    D = 256
    assert D in losses.SPLIT_DIMS
    words, regions, mask = word_inputs(B, D, T_, R, seed=B * R)
    g = torch.Generator().manual_seed(B)
    sent = planted_sent(B, 64, g)
    T.cfg.TRAIN.SMOOTH.GLOBAL = smooth
    try:
        labels = T.make_labels(B, sent.cuda(), b_global)
        r = regions.cuda().requires_grad_()
        w = words.cuda().requires_grad_()
        loss = T.word_loss(r, w, mask.cuda(), labels, b_global, precision="fp32")
        (1.3 * loss).backward()
        lab_o = oracle.make_labels(B, sent, b_global, smooth_global=smooth)
        ro, wo = regions.double().requires_grad_(), words.double().requires_grad_()
        lo = oracle.word_loss(ro, wo, mask, lab_o, b_global, smooth)
        (1.3 * lo).backward()
        # the CUDA-core form stays reachable
        r2, w2 = regions.cuda().requires_grad_(), words.cuda().requires_grad_()
        l2 = T.word_loss(r2, w2, mask.cuda(), labels, b_global, precision="fp32-simt")
        (1.3 * l2).backward()
    finally:
        T.cfg.TRAIN.SMOOTH.GLOBAL = 0.5
    assert lerr(loss.detach(), lo.detach()) <= TOL_FP32
    assert nerr(r.grad, ro.grad) <= TOL_FP32 and nerr(w.grad, wo.grad) <= TOL_FP32, (nerr(r.grad, ro.grad), nerr(w.grad, wo.grad))
    assert lerr(l2.detach(), lo.detach()) <= TOL_FP32 and nerr(r2.grad, ro.grad) <= TOL_FP32 and nerr(w2.grad, wo.grad) <= TOL_FP32


def test_full_size_fp32_mode_properties():
    """COCO-256 in the fp32 mode (BASELINE config 2): finite, padding words get exactly zero gradient, the loss of a
    batch whose captions match their own images' regions is below the chance level 2 ln B."""
    import math
    import bench
    from xmc_gan_b200 import train_gan as T
    inp = bench.make_inputs(256, 3)
    r = inp["regions"].cuda().requires_grad_()
    w = inp["words"].cuda().requires_grad_()
    m = inp["mask"].cuda()
    loss = T.word_loss(r, w, m, T.make_labels(256, None, False), False, rho1=5.0, rho2=5.0, rho3=10.0, precision="fp32")
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(r.grad).all() and torch.isfinite(w.grad).all()
    assert float(w.grad.transpose(1, 2)[m].abs().max()) == 0.0
    assert float(loss) < 2 * math.log(256)
