"""Parity of the layout prologue/epilogue of the word loss (xmc_normalize_transpose and its backward):
against a plain torch restatement of F.normalize + transpose (train_gan.py:88-89 convention), and the
bf16 fast kernels against the generic ones, bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from xmc_gan_b200.ops import default_ops
    return default_ops()


def _hops():
    """The hooks build: the only one in which the generic-kernel switch exists."""
    from util import hooks_ops
    return hooks_ops()


def _generic(on):
    _hops().L.xmc_internal_set_prep_generic(int(on))


def _reference(x, Lpad, row_of=None):
    """fp64 restatement: unit rows [B, Lpad, D] (zero rows beyond L) and the norms [B, Lpad]."""
    B, D, L = x.shape
    x = x.double()
    n = x.norm(dim=1).clamp_min(1e-12)
    xn = torch.zeros(B, Lpad, D, dtype=torch.float64, device=x.device)
    xn[:, :L] = (x / n[:, None]).transpose(1, 2)
    norm = torch.zeros(B, Lpad, dtype=torch.float64, device=x.device)
    norm[:, :L] = n
    if row_of is not None:
        keep = row_of >= 0
        out = torch.zeros_like(xn)
        out.view(B * L, D)[row_of[keep].long()] = xn.view(B * L, D)[keep]
        xn = out
    return xn, norm


def _rows(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] >= lens[:, None]).to(torch.uint8).cuda()
    return _ops().word_rows_compact(mask)[0]


CASES = [  # B, D, L, Lpad
    (3, 256, 289, 304), (2, 128, 289, 304), (4, 256, 18, 18), (2, 128, 33, 48),
    (2, 64, 40, 48), (2, 48, 7, 16), (1, 512, 65, 80),
]


@pytest.mark.parametrize("B,D,L,Lpad", CASES)
@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                                (torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_normalize_transpose_matches_torch(B, D, L, Lpad, in_dtype, out_dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + D + L)
    x = (torch.randn(B, D, L, generator=g) * torch.rand(B, 1, L, generator=g) * 3).to(in_dtype).cuda()
    x[0, :, L // 2] = 0                                   # a zero column: norm clamps to eps, unit row stays zero
    xn, norm = ops.normalize_transpose(x, Lpad, out_dtype)
    ref_xn, ref_n = _reference(x, Lpad)
    tol = 1e-6 if out_dtype == torch.float32 else 4e-3     # bf16 rounding of values <= 1: 2^-9
    assert (xn.double() - ref_xn).abs().max() <= tol
    assert torch.allclose(norm.double(), ref_n, rtol=1e-6, atol=1e-12)
    assert xn[:, L:].abs().max() == 0 if Lpad > L else True


@pytest.mark.parametrize("B,D,L,Lpad", CASES)
@pytest.mark.parametrize("xn_dtype,out_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                                (torch.bfloat16, torch.float32)])
def test_normalize_transpose_backward_matches_autograd(B, D, L, Lpad, xn_dtype, out_dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(B + D * 7 + L)
    x = torch.randn(B, D, L, generator=g).to(xn_dtype).cuda()
    xn, norm = ops.normalize_transpose(x, Lpad, xn_dtype)
    dxn = torch.randn(B, Lpad, D, generator=g).cuda()
    dnorm = torch.randn(B, Lpad, generator=g).cuda()
    dx = ops.normalize_transpose_backward(xn, norm, dxn, dnorm, L, out_dtype)
    # autograd through the fp64 restatement, evaluated at the (rounded) operands the kernel saw
    xl = x.double().requires_grad_()
    n = xl.norm(dim=1)
    y = (xl / n[:, None]).transpose(1, 2)
    ((y * dxn[:, :L].double()).sum() + (n * dnorm[:, :L].double()).sum()).backward()
    err = float((dx.double() - xl.grad).norm() / xl.grad.norm())
    assert err < (2e-6 if (xn_dtype, out_dtype) == (torch.float32, torch.float32) else 6e-3), err


@pytest.mark.parametrize("B,D,T", [(16, 256, 18), (5, 128, 18), (7, 256, 33)])
def test_compact_rows_forward_and_backward(B, D, T):
    ops = _ops()
    g = torch.Generator().manual_seed(D + T)
    row_of = _rows(B, T, 3)
    x = torch.randn(B, D, T, generator=g).bfloat16().cuda()
    xn, norm = ops.normalize_transpose(x, T, torch.bfloat16, row_of=row_of)
    ref_xn, _ = _reference(x, T, row_of)
    assert (xn.double() - ref_xn).abs().max() <= 4e-3
    dxn = torch.randn(B, T, D, generator=g).cuda()
    dx = ops.normalize_transpose_backward(xn, norm, dxn, None, T, torch.float32, row_of=row_of)
    # dense restatement: gather the compact gradient rows back, zero for dropped words
    keep = (row_of >= 0).view(B, T)
    dense_g = torch.zeros(B * T, D, device="cuda")
    dense_g[keep.flatten()] = dxn.view(B * T, D)[row_of[row_of >= 0].long()]
    xd, nd = ops.normalize_transpose(x, T, torch.bfloat16)
    ref = ops.normalize_transpose_backward(xd, nd, dense_g.view(B, T, D), None, T, torch.float32)
    ref = ref * keep[:, None, :]
    assert torch.equal(dx, ref)


@pytest.mark.parametrize("B,D,L,Lpad,compact", [(3, 256, 289, 304, False), (2, 128, 289, 304, False),
                                                (16, 256, 18, 18, True), (5, 128, 33, 33, True), (2, 256, 31, 32, False)])
def test_bf16_fast_kernels_are_bit_identical_to_generic(B, D, L, Lpad, compact):
    ops = _hops()                                     # generic = hooks build with the switch on; fast = the same build, switch off
    prod = _ops()
    g = torch.Generator().manual_seed(L)
    x = torch.randn(B, D, L, generator=g).bfloat16().cuda()
    x[0, :, 1] = 0
    row_of = _rows(B, L, 5) if compact else None
    dxn = torch.randn(B, Lpad, D, generator=g).cuda()
    dnorm = None if compact else torch.randn(B, Lpad, generator=g).cuda()
    out = {}
    try:
        for generic in (True, False):
            _generic(generic)
            xn, norm = ops.normalize_transpose(x, Lpad, torch.bfloat16, row_of=row_of)
            d16 = ops.normalize_transpose_backward(xn, norm, dxn, dnorm, L, torch.bfloat16, row_of=row_of)
            d32 = ops.normalize_transpose_backward(xn, norm, dxn, dnorm, L, torch.float32, row_of=row_of)
            out[generic] = (xn, norm, d16, d32)
    finally:
        _generic(False)
    for a, b in zip(out[True], out[False]):
        assert torch.equal(a, b)
    # and the product library (no switch at all) takes the fast kernels: same bits again
    xn, norm = prod.normalize_transpose(x, Lpad, torch.bfloat16, row_of=row_of)
    d16 = prod.normalize_transpose_backward(xn, norm, dxn, dnorm, L, torch.bfloat16, row_of=row_of)
    for a, b in zip((xn, norm, d16), out[False][:3]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("dense_labels,divs", [(False, False), (True, True)])
def test_fused_word_tail_backward_equals_the_two_calls(compact, dense_labels, divs):
    """xmc_word_scores_infonce_backward == xmc_infonce_grad followed by xmc_word_scores_backward, bit for bit."""
    ops = _ops()
    Bi, Bc, T = 37, 29, 11
    g = torch.Generator().manual_seed(7)
    lens = torch.randint(0, T + 1, (Bc,), generator=g)
    mask = (torch.arange(T)[None] >= lens[:, None]).to(torch.uint8).cuda()
    cap_ptr = None
    if compact:
        _, cap_ptr = ops.word_rows_compact(mask)
    rel = (torch.rand(Bi, Bc * T, generator=g) * 2 - 1).cuda()
    rho2, rho3 = 5.0, 10.0
    scores = ops.word_scores(rel, mask, Bc, T, rho2, cap_ptr=cap_ptr)
    labels = None
    if dense_labels:
        labels = (torch.rand(Bi, Bc, generator=g) > 0.8).float().cuda().contiguous()
    row_stats, col_stats = ops.infonce_stats(scores, labels, 3, rho3)
    row_div = (torch.randint(1, 4, (Bi,), generator=g).float().cuda()) if divs else None
    col_div = (torch.randint(1, 4, (Bc,), generator=g).float().cuda()) if divs else None
    go = torch.tensor(0.7, device="cuda")
    ds = ops.infonce_grad(scores, labels, 3, rho3, row_stats, col_stats, row_div, col_div, 1.0, Bi, Bc, go)
    two = ops.word_scores_backward(rel, mask, scores, ds, T, rho2, cap_ptr=cap_ptr)
    one = ops.word_scores_infonce_backward(rel, mask, scores, T, rho2, labels, 3, rho3, row_stats, col_stats,
                                           row_div, col_div, 1.0, Bi, Bc, go, cap_ptr=cap_ptr)
    n = int(cap_ptr[Bc]) if compact else Bc * T          # compact: columns beyond the valid rows are never written
    assert torch.equal(one[:, :n], two[:, :n])
    assert two[:, :n].abs().sum() > 0


@pytest.mark.parametrize("world,Bk", [(2, 512), (8, 2048), (3, 37)])
def test_combine_col_stats_matches_torch(world, Bk):
    """xmc_infonce_combine_stats == logsumexp over shards (row 0) and sums (rows 1-2)."""
    ops = _ops()
    g = torch.Generator().manual_seed(world * 1000 + Bk)
    gathered = (torch.randn(world, 3, Bk, generator=g) * 5).cuda()
    gathered[:, 0, 3] = float("-inf")                      # a column no shard has seen
    gathered[0, 0, 5] = float("-inf")                      # ... and one a single shard has not
    out = ops.combine_col_stats(gathered)
    ref0 = torch.logsumexp(gathered[:, 0].double(), dim=0)
    assert torch.equal(torch.isinf(out[0]), torch.isinf(ref0)) and bool((out[0][torch.isinf(out[0])] < 0).all())
    fin = ~torch.isinf(ref0)
    assert torch.allclose(out[0][fin].double(), ref0[fin], rtol=1e-6, atol=1e-6)
    assert torch.allclose(out[1:].double(), gathered[:, 1:].double().sum(0), rtol=1e-6, atol=1e-5)


@pytest.mark.parametrize("B,D,H,W", [(3, 256, 16, 16), (2, 256, 17, 17), (4, 128, 8, 8), (2, 64, 5, 3), (2, 512, 4, 6)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_row_layout_kernels_equal_the_transposing_ones(B, D, H, W, dt):
    """xmc_normalize_rows (+ backward) on a channels-last map == xmc_normalize_transpose (+ backward) on the
    contiguous map: same arithmetic, no transpose."""
    ops = _ops()
    g = torch.Generator().manual_seed(B * D + H)
    x = torch.randn(B, D, H, W, generator=g).to(dt).cuda()
    x[0, :, 0, 1] = 0
    L, Lpad = H * W, (H * W + 15) // 16 * 16
    rows = x.permute(0, 2, 3, 1).contiguous().view(B, L, D)
    xn_r, n_r = ops.normalize_rows(rows, Lpad, dt)
    xn_t, n_t = ops.normalize_transpose(x.flatten(2).contiguous(), Lpad, dt)
    assert torch.allclose(n_r, n_t, rtol=1e-6, atol=0)                      # same sum, different order
    assert float((xn_r.float() - xn_t.float()).abs().max()) <= (2 ** -8 if dt == torch.bfloat16 else 1e-6)   # <= 1 bf16 ulp of a unit row
    dxn = torch.randn(B, Lpad, D, generator=g).cuda()
    dnorm = torch.randn(B, Lpad, generator=g).cuda()
    d_r = ops.normalize_rows_backward(xn_t, n_t, dxn, dnorm, L, torch.float32)                     # [B, L, D]
    d_t = ops.normalize_transpose_backward(xn_t, n_t, dxn, dnorm, L, torch.float32)               # [B, D, L]
    err = float((d_r.transpose(1, 2) - d_t).norm() / d_t.norm())
    assert err <= 1e-6, err


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("fp32", 1e-4)])
def test_word_loss_on_channels_last_regions(precision, tol):
    """The word loss fed a channels-last region map (what a 1x1 region head writes): same loss and gradients as
    the contiguous map against the oracle, gradient returned channels-last."""
    import oracle
    from util import lerr, nerr, word_inputs
    from xmc_gan_b200 import train_gan as T
    B, D, T_, H = 12, 256, 9, 16
    words, regions, mask = word_inputs(B, D, T_, H * H, seed=3)
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    rd = (lambda v: v.to(dt).double())
    r4 = regions.view(B, D, H, H).to(dt).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    w = words.to(dt).cuda().requires_grad_()
    labels = T.make_labels(B, None, False)
    loss = T.word_loss(r4, w, mask.cuda(), labels, False, precision=precision)
    loss.backward()
    assert r4.grad.is_contiguous(memory_format=torch.channels_last)
    ro, wo = rd(regions).requires_grad_(), rd(words).requires_grad_()
    lo = oracle.word_loss(ro, wo, mask, torch.eye(B), False)
    lo.backward()
    assert lerr(loss.detach(), lo.detach()) <= tol
    assert nerr(r4.grad.flatten(2), ro.grad) <= tol and nerr(w.grad, wo.grad) <= tol
