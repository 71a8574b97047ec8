"""Region head fused into the word loss's prologue: the three tcgen05 products of
``xmc_gan_b200/csrc/region_head.cu`` through the C ABI against plain fp64 PyTorch on the bf16-rounded operands.

Forward: kn / rnorm of y = conv1x1(feat) + bias; backward: dfeat, dweight, dbias of a given dy.  bf16 operands are
exact inputs of the bf16 MMAs (the CPU side uses the same values: only accumulation order differs, 1e-3 of the
tensor's scale); an fp32 pair runs as tf32 MMAs (operands truncated to 10 mantissa bits: 2e-3); a mixed pair is rounded to
bf16 in registers (generic path; compared against the rounded values).  Unit rows are bf16: half an ulp of 1 on top."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from xmc_gan_b200.ops import default_ops
    return default_ops()


def _r(x):   # bf16-rounded values
    return x.bfloat16().double()


def _seen(a, b):
    """What the kernel multiplies: an fp32 pair with aligned rows stays fp32 (tf32 MMAs), anything else is bf16."""
    both_f32 = a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape[-1] % 4 == 0 and b.shape[-1] % 4 == 0
    return (a.double(), b.double(), 2e-3) if both_f32 else (_r(a), _r(b), 1e-3)


CASES = [  # B, Cin, H, W, feat dtype, weight dtype, bias
    (4, 512, 16, 16, torch.float32, torch.float32, True),
    (3, 512, 16, 16, torch.bfloat16, torch.bfloat16, True),
    (5, 512, 17, 17, torch.float32, torch.float32, True),      # R = 289: unaligned rows, three pixel tiles, pad rows
    (2, 192, 8, 8, torch.float32, torch.bfloat16, False),      # Cin not a multiple of the tiles, R < one tile
    (37, 256, 16, 16, torch.bfloat16, torch.float32, True),
    (1, 8, 1, 1, torch.float32, torch.float32, True),           # one pixel, one k-block mostly empty
    (3, 40, 4, 5, torch.bfloat16, torch.bfloat16, False),      # R = 20: rows not 16-byte aligned in bf16 -> generic path
    (2, 1024, 32, 32, torch.bfloat16, torch.bfloat16, True),   # R = 1024: four pixel tiles per image, dfeat in four column tiles
]


@pytest.mark.parametrize("B,Cin,H,W,fdt,wdt,has_bias", CASES)
def test_forward_matches_conv_plus_normalize(B, Cin, H, W, fdt, wdt, has_bias):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + Cin)
    D, R = 256, H * W
    feat = torch.randn(B, Cin, R, generator=g).to(fdt)
    w = (torch.randn(D, Cin, generator=g) / Cin ** 0.5).to(wdt)
    bias = torch.randn(D, generator=g) * 0.1 if has_bias else None
    Rpad = (R + 15) // 16 * 16
    kn, rnorm = ops.region_head_forward(feat.cuda(), w.cuda(), None if bias is None else bias.cuda(), Rpad)
    fs, ws, tol = _seen(feat, w)
    y = torch.einsum("bcr,dc->brd", fs, ws) + (0 if bias is None else bias.double())
    n = y.norm(dim=-1)
    assert kn.shape == (B, Rpad, D) and kn.dtype == torch.bfloat16
    err_n = float((rnorm[:, :R].double().cpu() - n).abs().max() / n.max())
    err_k = float((kn[:, :R].double().cpu() - y / n.unsqueeze(-1)).abs().max())
    assert err_n < tol, err_n
    assert err_k < 6e-3, err_k            # bf16 resolution of a unit row's entries (|x| <= 1: half an ulp = 2^-9)
    if Rpad > R:
        assert float(kn[:, R:].float().abs().max()) == 0.0 and float(rnorm[:, R:].abs().max()) == 0.0


@pytest.mark.parametrize("B,Cin,H,W,fdt,wdt,has_bias", CASES)
def test_backward_matches_einsum(B, Cin, H, W, fdt, wdt, has_bias):
    ops = _ops()
    g = torch.Generator().manual_seed(B * 77 + Cin)
    D, R = 256, H * W
    feat = torch.randn(B, Cin, R, generator=g).to(fdt)
    w = (torch.randn(D, Cin, generator=g) / Cin ** 0.5).to(wdt)
    dy = (torch.randn(B, R, D, generator=g) * 0.01).to(fdt)          # as the loss hands it over: in the map's dtype
    dfeat, dw, db = ops.region_head_backward(feat.cuda(), w.cuda(), dy.cuda(), True, True, has_bias)
    wa, dya, tol_i = _seen(w, dy)
    fa, dyb, tol_w = _seen(feat, dy)
    ref_f = torch.einsum("brd,dc->bcr", dya, wa)
    ref_w = torch.einsum("brd,bcr->dc", dyb, fa)
    ref_b = dy.double().sum((0, 1))
    assert dfeat.dtype == fdt and dfeat.shape == feat.shape
    tol_f = tol_i if fdt == torch.float32 else 6e-3                # bf16 output: its own rounding
    assert float((dfeat.double().cpu() - ref_f).abs().max() / ref_f.abs().max()) < tol_f
    assert float((dw.double().cpu() - ref_w).abs().max() / ref_w.abs().max()) < tol_w
    if has_bias:
        assert float((db.double().cpu() - ref_b).abs().max() / ref_b.abs().max()) < 1e-3
    else:
        assert db is None


def test_full_size_head_against_torch_on_gpu():
    """B = 256, [512, 16, 16] map (BASELINE config 4's discriminator stage): forward against conv + normalize in fp32 on the GPU."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, Cin, R, D = 256, 512, 256, 256
    feat = torch.randn(B, Cin, R, generator=g, device="cuda")
    w = torch.randn(D, Cin, generator=g, device="cuda") / Cin ** 0.5
    bias = torch.randn(D, generator=g, device="cuda") * 0.1
    kn, rnorm = ops.region_head_forward(feat, w, bias, R)
    y = torch.einsum("bcr,dc->brd", feat.double(), w.double()).float() + bias
    n = y.norm(dim=-1)
    assert float((rnorm - n).abs().max() / n.max()) < 2e-3
    assert float((kn.float() - y / n.unsqueeze(-1)).abs().max()) < 8e-3


def _head_inputs(B, Cin, H, W, T_, seed):
    """Feature map whose projection leans towards the caption's words (so the attention is not uniform)."""
    g = torch.Generator().manual_seed(seed)
    D = 256
    words = torch.randn(B, D, T_, generator=g)
    w = torch.randn(D, Cin, generator=g) / Cin ** 0.5
    bias = torch.randn(D, generator=g) * 0.05
    feat = torch.randn(B, Cin, H, W, generator=g)
    # add a component that projects onto a random word of the own caption: feat += W^+ e_t (least squares through W^T)
    pick = torch.randint(0, T_, (H * W,), generator=g)
    target = words[:, :, pick]                                        # [B, D, R]
    feat = feat + 0.5 * torch.einsum("dc,bdr->bcr", w, target).view(B, Cin, H, W)
    lens = torch.randint(max(1, T_ // 3), T_ + 1, (B,), generator=g)
    mask = torch.arange(T_).unsqueeze(0) >= lens.unsqueeze(1)
    return feat, w, bias, words, mask


@pytest.mark.parametrize("B,Cin,H,W,T_,fdt", [(16, 512, 16, 16, 18, torch.float32), (24, 512, 16, 16, 12, torch.bfloat16),
                                             (6, 256, 17, 17, 18, torch.float32)])
def test_word_loss_with_fused_head_vs_oracle(B, Cin, H, W, T_, fdt):
    """word_loss(feature map, region_head=(W, b), precision='bf16') == oracle.word_loss(conv1x1(map)) on the bf16-rounded
    operands: loss and the gradients of the map, the head's weight and bias, and the words (rel 2e-2)."""
    import oracle
    from util import TOL_BF16, lerr, nerr
    from xmc_gan_b200 import train_gan as T
    feat, w, bias, words, mask = _head_inputs(B, Cin, H, W, T_, seed=B + Cin)
    feat = feat.to(fdt)
    labels = T.make_labels(B, None, False)
    f = feat.clone().cuda().requires_grad_()
    wg = w.clone().cuda().view(256, Cin, 1, 1).requires_grad_()       # a Conv2d weight
    bg = bias.clone().cuda().requires_grad_()
    wd = words.bfloat16().cuda().requires_grad_()
    loss = T.word_loss(f, wd, mask.cuda(), labels, False, precision="bf16", region_head=(wg, bg))
    loss.backward()
    fo, wo, bo = _r(feat).requires_grad_(), _r(w).requires_grad_(), bias.double().requires_grad_()
    wdo = _r(words).requires_grad_()
    y = torch.einsum("bchw,dc->bdhw", fo, wo) + bo.view(1, -1, 1, 1)
    lo = oracle.word_loss(y, wdo, mask, torch.eye(B), False)
    lo.backward()
    assert lerr(loss, lo) <= TOL_BF16, (float(loss), float(lo))
    assert f.grad.dtype == fdt and wg.grad.shape == wg.shape
    for name, a, b in (("feat", f.grad, fo.grad), ("weight", wg.grad.view(256, Cin), wo.grad), ("bias", bg.grad, bo.grad),
                       ("words", wd.grad, wdo.grad)):
        assert nerr(a, b) <= TOL_BF16, (name, nerr(a, b))


def test_fused_head_equals_unfused_projection():
    """Same loss as projecting with PyTorch's convolution first (the fp32 fallback of region_head=) within the bf16 tolerance,
    through contrastive_losses as well."""
    from util import TOL_BF16, lerr, nerr
    from xmc_gan_b200 import train_gan as T
    B, Cin, H, W, T_ = 32, 512, 16, 16, 18
    feat, w, bias, words, mask = _head_inputs(B, Cin, H, W, T_, seed=5)
    labels = T.make_labels(B, None, False)
    conv = torch.nn.Conv2d(Cin, 256, 1).cuda()
    with torch.no_grad():
        conv.weight.copy_(w.view(256, Cin, 1, 1)); conv.bias.copy_(bias)
    f1 = feat.clone().cuda().requires_grad_()
    w1 = words.clone().cuda().requires_grad_()
    l1 = T.contrastive_losses(regions=f1, words=w1, mask=mask.cuda(), labels=labels, b_global=False, precision="bf16",
                              region_head=conv)[2]
    l1.backward()
    g_w, g_b = conv.weight.grad.clone(), conv.bias.grad.clone()
    conv.zero_grad()
    f2 = feat.clone().cuda().requires_grad_()
    w2 = words.clone().cuda().requires_grad_()
    l2 = T.word_loss(conv(f2), w2, mask.cuda(), labels, False, precision="bf16")
    l2.backward()
    assert lerr(l1, l2) <= TOL_BF16
    assert nerr(f1.grad, f2.grad) <= TOL_BF16 and nerr(w1.grad, w2.grad) <= TOL_BF16
    assert nerr(g_w, conv.weight.grad) <= TOL_BF16 and nerr(g_b, conv.bias.grad) <= TOL_BF16
    # fp32 precision: region_head= falls back to the convolution in front of the loss (fp32 tolerance path)
    f3 = feat.clone().cuda().requires_grad_()
    l3 = T.word_loss(f3, words.clone().cuda(), mask.cuda(), labels, False, precision="fp32", region_head=conv)
    l4 = T.word_loss(conv(f3), words.clone().cuda(), mask.cuda(), labels, False, precision="fp32")
    assert lerr(l3, l4) <= 1e-6


@pytest.mark.parametrize("B,C,H,dt", [(256, 512, 4, torch.float32), (7, 96, 4, torch.bfloat16), (5, 33, 3, torch.float32)])
def test_pooled_features_match_avg_pool2d(B, C, H, dt):
    """train_gan.pooled_features == F.avg_pool2d(x, H).view(B, -1) (df_gan.py:165-166, train_gan.py:271-276), forward and backward."""
    import torch.nn.functional as F
    from xmc_gan_b200 import train_gan as T
    g = torch.Generator().manual_seed(B + C)
    x0 = torch.randn(B, C, H, H, generator=g).to(dt)
    go = torch.randn(B, C, generator=g).to(dt)
    x = x0.clone().cuda().requires_grad_()
    y = T.pooled_features(x)
    y.backward(go.cuda())
    xr = x0.clone().cuda().requires_grad_()
    yr = F.avg_pool2d(xr, kernel_size=H).view(B, -1)
    yr.backward(go.cuda())
    tol = 1e-6 if dt == torch.float32 else 8e-3
    assert y.dtype == dt and y.shape == (B, C)
    assert float((y.float() - yr.float()).abs().max()) <= tol * max(1.0, float(yr.float().abs().max()))
    assert float((x.grad.float() - xr.grad.float()).abs().max()) <= tol * max(1.0, float(xr.grad.float().abs().max()))
    yb = T.pooled_features(x.detach(), out_dtype=torch.bfloat16)
    assert yb.dtype == torch.bfloat16 and float((yb.float() - yr.float()).abs().max()) <= 8e-3 * max(1.0, float(yr.float().abs().max()))


def test_cta_pair_variant_matches(monkeypatch):
    """Hooks build, XMC_HEAD_PAIR=1: the 2-SM form of the kernel (thread-block clusters of two, one tcgen05.mma.cta_group::2
    per 256-row tile, each CTA staging its own A rows and half of the B rows) gives the same forward and dfeat as single CTAs."""
    from util import hooks_ops
    ops = hooks_ops()                                     # the pair form is compiled into libxmcloss_hooks.so only
    g = torch.Generator(device="cuda").manual_seed(11)
    B, Cin, R, D = 6, 512, 256, 256
    for dt in (torch.bfloat16, torch.float32):
        feat = torch.randn(B, Cin, R, generator=g, device="cuda").to(dt)
        w = (torch.randn(D, Cin, generator=g, device="cuda") / Cin ** 0.5).to(dt)
        bias = torch.randn(D, generator=g, device="cuda") * 0.1
        dy = (torch.randn(B, R, D, generator=g, device="cuda") * 0.01).to(dt)
        monkeypatch.delenv("XMC_HEAD_PAIR", raising=False)
        kn0, rn0 = ops.region_head_forward(feat, w, bias, R)
        df0, _, _ = ops.region_head_backward(feat, w, dy, True, False, False)
        monkeypatch.setenv("XMC_HEAD_PAIR", "1")
        kn1, rn1 = ops.region_head_forward(feat, w, bias, R)
        df1, _, _ = ops.region_head_backward(feat, w, dy, True, False, False)
        torch.cuda.synchronize()
        assert torch.equal(kn0, kn1) and torch.equal(rn0, rn1)          # same products in the same order
        assert torch.equal(df0, df1)


def test_word_loss_with_fused_head_is_cuda_graph_capturable():
    """The fused head (TMA maps built per call on the host, memsets of the weight gradient, column-sum kernel) inside
    torch.cuda.graph: nothing synchronises with the host; replays reproduce the eager step."""
    from xmc_gan_b200 import train_gan as T
    B, Cin, H, W, T_ = 16, 512, 16, 16, 9
    feat, w, bias, words, mask = _head_inputs(B, Cin, H, W, T_, seed=21)
    f = feat.cuda().requires_grad_()
    wg = w.cuda().view(256, Cin, 1, 1).requires_grad_()
    bg = bias.cuda().requires_grad_()
    wd = words.bfloat16().cuda().requires_grad_()
    m = mask.cuda()
    labels = T.make_labels(B, None, False)
    leaves = (f, wg, bg, wd)

    def step():
        for t in leaves:
            t.grad = None
        loss = T.word_loss(f, wd, m, labels, False, precision="bf16", region_head=(wg, bg))
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
    ref = [t.grad.clone() for t in leaves]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    for _ in range(3):
        for t in leaves:
            t.grad.zero_()                   # the captured backward accumulates into the captured .grad tensors
        g.replay()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - ref_loss) <= 1e-5 * abs(ref_loss)
    for t, r in zip(leaves, ref):            # fp32 atomics: summation order differs from run to run
        assert float((t.grad.float() - r.float()).norm() / r.float().norm()) < 1e-2
