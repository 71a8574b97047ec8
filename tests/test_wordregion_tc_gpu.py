"""GPU tests of the tcgen05/TMEM/TMA word-region kernels (bf16 operands, fp32 accumulate).

Level 1: raw tensor-core plumbing — the TMEM debug dump of tile 0 (S = Q Khat^T of chunk 0 and the
context accumulator C) against torch matmuls on the same bf16 operands: this isolates descriptor /
swizzle / TMEM-layout mistakes from the loss maths.
Level 2: kernel statistics (lsum, cnorm, rel) against the fp32 CUDA-core kernel on the same operands.
Level 3: word_loss(precision='bf16') against the CPU oracle fed the bf16-rounded inputs, rel 2e-2."""
import ctypes

import pytest
import torch

import oracle
from util import TOL_BF16, lerr, nerr, word_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from xmc_gan_b200.ops import default_ops
    return default_ops()


def _operands(ops, B, D, T_, R, seed):
    words, regions, mask = word_inputs(B, D, T_, R, seed)
    Rpad = (R + 15) // 16 * 16
    qn, _ = ops.normalize_transpose(words.cuda(), T_, torch.bfloat16)
    kn, rnorm = ops.normalize_transpose(regions.cuda(), Rpad, torch.bfloat16)
    return qn.view(B * T_, D), kn, rnorm, mask


def _err_word(ops):
    torch.cuda.synchronize()
    return int(ops.last_workspace[:4].view(torch.int32)[0])


@pytest.mark.parametrize("D", [256, 128, 64])
def test_tmem_dump_matches_matmul(D):
    from util import hooks_ops
    from xmc_gan_b200 import _lib
    ops = hooks_ops()                                      # the debug dump exists only in the -DXMC_TEST_HOOKS build
    B, T_, R = 8, 18, 100
    qn, kn, rnorm, _ = _operands(ops, B, D, T_, R, seed=D)
    hook = ops.L.xmc_internal_set_debug_dump
    hook(1)
    try:
        lsum, cnorm, rel, _ = ops.wordregion_forward(_lib.PATH_BF16_TCGEN05, qn, kn, rnorm, R, 5.0)
        assert _err_word(ops) == 0, "mbarrier wait timed out inside the kernel"
        dump = ops.last_workspace[64:].view(torch.float32)
    finally:
        hook(0)
    q = qn[:128].float()                                   # tile 0 (zero rows beyond NQ)
    if q.shape[0] < 128:
        q = torch.cat([q, torch.zeros(128 - q.shape[0], D, device="cuda")])
    k0 = kn[0].float()                                     # image 0, [Rpad, D]
    S_ref = q @ k0[:64].t()
    S = dump[:128 * 64].view(128, 64)
    assert torch.allclose(S, S_ref, atol=2e-3), float((S - S_ref).abs().max())
    s_all = q @ k0.t()
    valid = (torch.arange(k0.shape[0], device="cuda") < R).float()
    p = torch.exp(5.0 * (s_all - 1.0)) * valid * rnorm[0]
    C_ref = p.bfloat16().float() @ k0
    C = dump[128 * 64:128 * 64 + 128 * D].view(128, D)
    assert nerr(C, C_ref) < 5e-3, nerr(C, C_ref)


@pytest.mark.parametrize("B,D,T_,R", [(8, 256, 18, 289), (32, 256, 18, 289), (5, 128, 7, 40), (16, 64, 32, 64),
                                      (3, 256, 12, 17), (40, 256, 20, 256)])
@pytest.mark.parametrize("raw_values", [True, False])
def test_statistics_vs_fp32_kernel(ops, B, D, T_, R, raw_values):
    from xmc_gan_b200 import _lib
    qn, kn, rnorm, _ = _operands(ops, B, D, T_, R, seed=B + R)
    rn = rnorm if raw_values else None
    l1, c1, r1, _ = ops.wordregion_forward(_lib.PATH_BF16_TCGEN05, qn, kn, rn, R, 5.0)
    assert _err_word(ops) == 0
    l0, c0, r0, _ = ops.wordregion_forward(_lib.PATH_FP32_SIMT, qn.float(), kn.float(), rn, R, 5.0)
    assert nerr(l1, l0) < 1e-3, nerr(l1, l0)
    assert nerr(c1, c0) < 1e-2, nerr(c1, c0)
    assert float((r1 - r0).abs().max()) < 2e-2, float((r1 - r0).abs().max())


@pytest.mark.parametrize("B,D,T_,R", [(8, 256, 18, 289), (32, 256, 18, 289), (5, 128, 7, 40), (3, 256, 12, 17),
                                      (40, 256, 20, 256), (20, 128, 32, 100)])
@pytest.mark.parametrize("raw_values", [True, False])
def test_backward_vs_fp32_kernel(ops, B, D, T_, R, raw_values):
    """tcgen05 backward (dQ, dKhat, d rnorm) against the fp32 CUDA-core backward on the same operands,
    statistics and upstream gradient."""
    from xmc_gan_b200 import _lib
    qn, kn, rnorm, _ = _operands(ops, B, D, T_, R, seed=7 * B + R)
    rn = rnorm if raw_values else None
    l1, c1, r1, chat = ops.wordregion_forward(_lib.PATH_BF16_TCGEN05, qn, kn, rn, R, 5.0, save_context=True)
    assert _err_word(ops) == 0
    g = torch.Generator().manual_seed(B)
    grel = torch.randn(l1.shape, generator=g).cuda() * 0.1
    dq1, dk1, dr1 = ops.wordregion_backward(_lib.PATH_BF16_TCGEN05, qn, kn, rn, R, 5.0, l1, c1, r1, grel, chat)
    assert _err_word(ops) == 0, "mbarrier wait timed out inside the backward kernel"
    dq0, dk0, dr0 = ops.wordregion_backward(_lib.PATH_FP32_SIMT, qn.float(), kn.float(), rn, R, 5.0, l1, c1, r1, grel)
    # context sums (unscaled softmax numerators times values) saved by the forward
    s_all = torch.einsum('qd,ird->iqr', qn.float(), kn.float())
    valid = (torch.arange(kn.shape[1], device="cuda") < R).float()
    pw = torch.exp(5.0 * (s_all - 1.0)) * valid * (rnorm.unsqueeze(1) if raw_values else 1.0)
    ctx = torch.einsum('iqr,ird->iqd', pw, kn.float())
    chat_ref = ctx
    assert nerr(chat, chat_ref) < 1e-2, nerr(chat, chat_ref)
    assert nerr(dq1, dq0) < 1.5e-2, nerr(dq1, dq0)
    assert nerr(dk1, dk0) < 1.5e-2, nerr(dk1, dk0)
    if raw_values:
        assert nerr(dr1, dr0) < 1.5e-2, nerr(dr1, dr0)


@pytest.mark.parametrize("B,D,T_,R", [(32, 256, 18, 289), (12, 128, 9, 64), (64, 256, 18, 289), (9, 64, 5, 30)])
def test_word_loss_bf16_vs_oracle(B, D, T_, R):
    from xmc_gan_b200 import train_gan as T
    words, regions, mask = word_inputs(B, D, T_, R, seed=3 * B)
    wb, rb = words.bfloat16(), regions.bfloat16()
    labels = T.make_labels(B, None, False)
    r = rb.clone().cuda().requires_grad_()
    w = wb.clone().cuda().requires_grad_()
    loss = T.word_loss(r, w, mask.cuda(), labels, False, precision="bf16")
    loss.backward()
    ro, wo = rb.double().requires_grad_(), wb.double().requires_grad_()
    lo = oracle.word_loss(ro, wo, mask, torch.eye(B), False)
    lo.backward()
    assert lerr(loss, lo) <= TOL_BF16, (float(loss), float(lo))
    assert nerr(r.grad, ro.grad) <= TOL_BF16, nerr(r.grad, ro.grad)
    assert nerr(w.grad, wo.grad) <= TOL_BF16, nerr(w.grad, wo.grad)


@pytest.mark.parametrize("Bc,T_", [(7, 18), (256, 18), (100, 32), (3, 5)])
def test_word_rows_compact(ops, Bc, T_):
    g = torch.Generator().manual_seed(Bc)
    mask = torch.rand(Bc, T_, generator=g) < 0.4          # arbitrary pattern, not only suffix padding
    mask[0] = True                                         # a fully padded caption
    row_of, cap_ptr = ops.word_rows_compact(mask.to(torch.uint8).cuda())
    valid = (~mask).flatten()
    ref = torch.where(valid, torch.cumsum(valid.int(), 0) - 1, torch.full_like(valid.int(), -1))
    assert torch.equal(row_of.cpu(), ref.int())
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), (~mask).sum(1).cumsum(0)])
    assert torch.equal(cap_ptr.cpu().long(), ptr)


@pytest.mark.parametrize("B,D,T_,R", [(32, 256, 18, 289), (12, 128, 9, 64), (150, 256, 18, 40)])
def test_compact_rows_match_dense_rows(B, D, T_, R):
    """The compacted schedule (valid word rows only, device-side count, persistent CTAs) and the dense one
    (every padded row visited) give the same loss and gradients."""
    from xmc_gan_b200 import train_gan as T
    from xmc_gan_b200.ops import default_ops
    words, regions, mask = word_inputs(B, D, T_, R, seed=11 * B)
    mask[1] = True                                         # one fully padded caption
    labels = T.make_labels(B, None, False)
    out = []
    ops = default_ops()
    for compact in (True, False):
        ops.supports_compaction = compact
        try:
            r = regions.bfloat16().cuda().requires_grad_()
            w = words.bfloat16().cuda().requires_grad_()
            loss = T.word_loss(r, w, mask.cuda(), labels, False, precision="bf16")
            loss.backward()
            out.append((loss.detach(), r.grad, w.grad))
        finally:
            ops.supports_compaction = True
    (l1, r1, w1), (l0, r0, w0) = out
    assert lerr(l1, l0) <= 1e-5, (float(l1), float(l0))
    assert nerr(r1, r0) <= 2e-3, nerr(r1, r0)             # bf16 outputs, different summation order
    assert nerr(w1, w0) <= 2e-3, nerr(w1, w0)
    assert float(w1[1].abs().max()) == 0.0                 # padded words: exactly zero gradient
    assert float((w1 * mask.cuda()[:, None, :]).abs().max()) == 0.0


def test_full_size_coco256_tc_matches_fp32_path():
    """BASELINE config 2 at full size (B=256, T=18, R=17x17, D=256, caption lengths U{5..18}): the tcgen05
    path (compacted rows, persistent schedule over 148 CTAs) against the fp32 CUDA-core path, which is itself
    checked against the oracle at the sizes the oracle can run.  Plus linearity in the upstream gradient."""
    from xmc_gan_b200 import train_gan as T
    B, D, T_, R = 256, 256, 18, 289
    words, regions, mask = word_inputs(B, D, T_, R, seed=2024, min_len=5)
    wb, rb = words.bfloat16(), regions.bfloat16()
    labels = T.make_labels(B, None, False)

    def run(precision, scale):
        r = rb.cuda().to(torch.float32 if precision == "fp32" else torch.bfloat16).requires_grad_()
        w = wb.cuda().to(torch.float32 if precision == "fp32" else torch.bfloat16).requires_grad_()
        loss = T.word_loss(r.view(B, D, 17, 17), w, mask.cuda(), labels, False, precision=precision)
        (loss * scale).backward()
        return loss.detach(), r.grad.float(), w.grad.float()

    l_tc, r_tc, w_tc = run("bf16", 1.0)
    l_32, r_32, w_32 = run("fp32", 1.0)
    assert torch.isfinite(l_tc) and torch.isfinite(r_tc).all() and torch.isfinite(w_tc).all()
    assert lerr(l_tc, l_32) <= TOL_BF16, (float(l_tc), float(l_32))
    assert nerr(r_tc, r_32) <= TOL_BF16, nerr(r_tc, r_32)
    assert nerr(w_tc, w_32) <= TOL_BF16, nerr(w_tc, w_32)
    _, r_3, w_3 = run("bf16", 3.0)                         # gradients are linear in grad_out
    assert nerr(r_3, 3.0 * r_tc) <= 6e-3, nerr(r_3, 3.0 * r_tc)      # bf16 outputs + atomic summation order
    assert nerr(w_3, 3.0 * w_tc) <= 6e-3, nerr(w_3, 3.0 * w_tc)
    assert float((w_tc * mask.cuda()[:, None, :]).abs().max()) == 0.0


def test_more_word_tiles_than_sms(ops):
    """Rows local / columns gathered, as one rank of the 8-GPU run sees it: 1100 captions x 18 words = 155 word
    tiles (> 148 SMs: full rounds plus a shared last round of the persistent schedule) against 40 images."""
    from xmc_gan_b200 import _lib
    Bc, Bi, D, T_, R = 1100, 40, 256, 18, 100
    g = torch.Generator().manual_seed(5)
    words = torch.randn(Bc, D, T_, generator=g)
    regions = torch.randn(Bi, D, R, generator=g) + 0.3 * words[:Bi, :, torch.randint(0, T_, (R,), generator=g)]
    lens = torch.randint(5, T_ + 1, (Bc,), generator=g)
    mask = (torch.arange(T_).unsqueeze(0) >= lens.unsqueeze(1)).to(torch.uint8).cuda()
    Rpad = (R + 15) // 16 * 16
    row_of, cap_ptr = ops.word_rows_compact(mask)
    nq = cap_ptr[Bc:]
    qc, _ = ops.normalize_transpose(words.cuda(), T_, torch.bfloat16, row_of=row_of)
    qd, _ = ops.normalize_transpose(words.cuda(), T_, torch.bfloat16)
    kn, rnorm = ops.normalize_transpose(regions.cuda(), Rpad, torch.bfloat16)
    qc, qd = qc.view(Bc * T_, D), qd.view(Bc * T_, D)
    l1, c1, r1, chat = ops.wordregion_forward(_lib.PATH_BF16_TCGEN05, qc, kn, rnorm, R, 5.0, save_context=True, nq_dev=nq)
    assert _err_word(ops) == 0
    l0, c0, r0, _ = ops.wordregion_forward(_lib.PATH_FP32_SIMT, qd.float(), kn.float(), rnorm, R, 5.0)
    valid = row_of >= 0
    idx = row_of[valid].long()
    assert nerr(r1[:, idx], r0[:, valid]) < 5e-3
    assert nerr(l1[:, idx], l0[:, valid]) < 5e-3
    grel_d = torch.randn(Bi, Bc * T_, generator=g).cuda() * 0.1 * valid.float()
    grel_c = torch.zeros_like(grel_d)
    grel_c[:, idx] = grel_d[:, valid]
    dq1, dk1, dr1 = ops.wordregion_backward(_lib.PATH_BF16_TCGEN05, qc, kn, rnorm, R, 5.0, l1, c1, r1, grel_c, chat, nq_dev=nq)
    assert _err_word(ops) == 0
    dq0, dk0, dr0 = ops.wordregion_backward(_lib.PATH_FP32_SIMT, qd.float(), kn.float(), rnorm, R, 5.0, l0, c0, r0, grel_d)
    assert nerr(dq1[idx], dq0[valid]) < 1.5e-2
    assert nerr(dk1, dk0) < 1.5e-2
    assert nerr(dr1, dr0) < 1.5e-2


@pytest.mark.parametrize("B,D,T_,R", [(6, 256, 18, 40), (5, 128, 7, 17), (9, 256, 12, 289), (4, 256, 18, 64)])
def test_never_written_tmem_columns_are_not_read(B, D, T_, R):
    """Chunks narrower than 64 regions leave TMEM columns the MMAs never write.  With tensor memory filled
    with NaNs first (debug flag 16) the result must not change: nothing stale may leak into the sums."""
    from util import hooks_ops
    from xmc_gan_b200 import train_gan as T
    hops = hooks_ops()                                     # the NaN-poisoning flag exists only in the hooks build
    hook = hops.L.xmc_internal_set_debug_dump
    words, regions, mask = word_inputs(B, D, T_, R, seed=B + R)
    labels = T.make_labels(B, None, False)
    out = []
    for flag in (0, 16):
        hook(flag)
        try:
            r = regions.bfloat16().cuda().requires_grad_()
            w = words.bfloat16().cuda().requires_grad_()
            loss = T.word_loss(r, w, mask.cuda(), labels, False, precision="bf16", _ops=hops)
            loss.backward()
            torch.cuda.synchronize()
            out.append((loss.detach(), r.grad.float(), w.grad.float()))
        finally:
            hook(0)
    (l0, r0, w0), (l1, r1, w1) = out
    assert torch.isfinite(l1) and torch.isfinite(r1).all() and torch.isfinite(w1).all()
    assert lerr(l1, l0) <= 1e-6 and nerr(r1, r0) <= 2e-3 and nerr(w1, w0) <= 2e-3


def test_word_loss_is_cuda_graph_capturable():
    """The whole word loss (side-stream prologue, device-side row count, fused tail backward) can be captured
    with torch.cuda.graph and replayed: nothing in it synchronises with the host, and the side stream it
    forks is joined before the autograd function returns.  Replays must reproduce the eager result."""
    from util import word_inputs
    from xmc_gan_b200 import train_gan as T
    B, D, T_, R = 24, 128, 9, 70
    words, regions, mask = word_inputs(B, D, T_, R, seed=11)
    r = regions.bfloat16().cuda().requires_grad_()
    w = words.bfloat16().cuda().requires_grad_()
    m = mask.cuda()
    labels = T.make_labels(B, torch.randn(B, 16).cuda(), False)

    def step():
        r.grad = None; w.grad = None
        loss = T.word_loss(r, w, m, labels, False, precision="bf16")
        loss.backward()
        return loss

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    ref_loss = float(step().detach())
    ref_r, ref_w = r.grad.clone(), w.grad.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    for _ in range(3):
        r.grad.zero_(); w.grad.zero_()       # the captured backward ACCUMULATES into the captured .grad tensors
        g.replay()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - ref_loss) <= 1e-5 * abs(ref_loss)
    # fp32 atomics: summation order differs from run to run
    assert float((r.grad.float() - ref_r.float()).norm() / ref_r.float().norm()) < 1e-2
    assert float((w.grad.float() - ref_w.float()).norm() / ref_w.float().norm()) < 1e-2


def test_attention_stage_of_the_kernels_against_the_reference_attention_block(ops):
    """The forward kernels' attention stage — cosines, softmax over the regions, contexts — against the REFERENCE's own
    attention code (xmc_gan/model/concept_gan.py:532-555, run on every (image, caption) pair by tests/golden/make_golden.py):
    rho1 = 1, unit values.  The tcgen05 kernel saves C = l * context (bf16) and ||context||; the fp32-tolerance kernel the
    same as a hi + lo pair.  This pins the CUDA path's attention on reference code, not on this repo's oracle."""
    import glob
    import os
    import numpy as np
    from xmc_gan_b200 import _lib
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_attn_*.npz"))):
        g = np.load(path)
        words, regions = torch.from_numpy(g["words"]), torch.from_numpy(g["regions"])
        ref = torch.from_numpy(g["ctx64"])                                  # [Bi, Bc, T, D]
        Bi, Bc, T_, D = ref.shape
        R = regions.shape[2]
        Rpad = (R + 15) // 16 * 16
        for path_id, dt, tol in ((_lib.PATH_BF16_TCGEN05, torch.bfloat16, 1.5e-2), (_lib.PATH_FP32_TCGEN05, torch.float32, 1e-4)):
            if path_id == _lib.PATH_FP32_TCGEN05 and D != 256:
                continue
            qn, _ = ops.normalize_transpose(words.cuda(), T_, dt)
            kn, _ = ops.normalize_transpose(regions.cuda(), Rpad, dt)
            lsum, cnorm, rel, chat = ops.wordregion_forward(path_id, qn.view(Bc * T_, D), kn, None, R, 1.0, save_context=True)
            torch.cuda.synchronize()
            c = chat.double().sum(0) if chat.dim() == 4 else chat.double()  # hi + lo planes
            ctx = (c / lsum.double().unsqueeze(-1)).view(Bi, Bc, T_, D).cpu()
            scale = float(ref.norm(dim=-1).mean())
            assert float((ctx - ref).norm(dim=-1).max()) <= tol * max(1.0, scale) , (path, path_id, float((ctx - ref).norm(dim=-1).max()))
            assert float((cnorm.double().view(Bi, Bc, T_).cpu() - ref.norm(dim=-1)).abs().max()) <= tol, (path, path_id)
