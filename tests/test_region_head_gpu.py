"""Region head fused into the word loss's prologue: the three tcgen05 products of
``xmc_gan_b200/csrc/region_head.cu`` through the C ABI against plain fp64 PyTorch on the bf16-rounded operands.

Forward: kn / rnorm of y = conv1x1(feat) + bias; backward: dfeat, dweight, dbias of a given dy.  Tolerances are the
bf16 mode's (operands rounded to bf16, fp32 accumulation): unit rows to bf16 resolution, norms and gradients 1e-3
relative to the tensor's scale (the CPU side uses the same rounded operands, so only accumulation order differs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from xmc_gan_b200.ops import default_ops
    return default_ops()


def _r(x):   # what the kernel sees: bf16-rounded values
    return x.bfloat16().double()


CASES = [  # B, Cin, H, W, feat dtype, weight dtype, bias
    (4, 512, 16, 16, torch.float32, torch.float32, True),
    (3, 512, 16, 16, torch.bfloat16, torch.bfloat16, True),
    (5, 512, 17, 17, torch.float32, torch.float32, True),      # R = 289: unaligned rows, three pixel tiles, pad rows
    (2, 192, 8, 8, torch.float32, torch.bfloat16, False),      # Cin not a multiple of the tiles, R < one tile
    (37, 256, 16, 16, torch.bfloat16, torch.float32, True),
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
    y = torch.einsum("bcr,dc->brd", _r(feat), _r(w)) + (0 if bias is None else bias.double())
    n = y.norm(dim=-1)
    assert kn.shape == (B, Rpad, D) and kn.dtype == torch.bfloat16
    err_n = float((rnorm[:, :R].double().cpu() - n).abs().max() / n.max())
    err_k = float((kn[:, :R].double().cpu() - y / n.unsqueeze(-1)).abs().max())
    assert err_n < 1e-4, err_n
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
    dy = (torch.randn(B, R, D, generator=g) * 0.01).bfloat16()
    dfeat, dw, db = ops.region_head_backward(feat.cuda(), w.cuda(), dy.cuda(), True, True, has_bias)
    ref_f = torch.einsum("brd,dc->bcr", dy.double(), _r(w))
    ref_w = torch.einsum("brd,bcr->dc", dy.double(), _r(feat))
    ref_b = dy.double().sum((0, 1))
    assert dfeat.dtype == fdt and dfeat.shape == feat.shape
    tol_f = 1e-3 if fdt == torch.float32 else 6e-3                 # bf16 output: its own rounding
    assert float((dfeat.double().cpu() - ref_f).abs().max() / ref_f.abs().max()) < tol_f
    assert float((dw.double().cpu() - ref_w).abs().max() / ref_w.abs().max()) < 1e-3
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
    y = torch.einsum("bcr,dc->brd", feat.bfloat16().float(), w.bfloat16().float()) + bias
    n = y.norm(dim=-1)
    assert float((rnorm - n).abs().max() / n.max()) < 1e-3
    assert float((kn.float() - y / n.unsqueeze(-1)).abs().max()) < 8e-3
