"""BASELINE config 4 / SURVEY §8f N3 on the GPU: the reference's G/D update (xmc_gan/train_gan.py:187-289, as
``xmc_gan_b200.step.gd_step``) with the loss ops swapped.  The same networks, inputs and noise go through the step
twice — once with the stock-PyTorch loss block (tests/stock_losses.py = the oracle's restatements on the GPU), once
with this package's kernels — and every parameter gradient of the discriminator update and of the generator update
must agree: fp32 rel <= 1e-4 (north_star), bf16 word loss rel <= 2e-2."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
pytestmark = pytest.mark.gpu


def _setup(B, size, nch, nef, T_, seed=0):
    from dfgan_harness import NetD, NetG
    torch.manual_seed(seed)
    G = NetG(size, nch, 100, nef, nef).cuda()
    D = NetD(size, nch, nef, img_match=True, spec_norm=False, region_res=16).cuda()
    for m in (G, D):
        for n, p in m.named_parameters():
            if n.endswith("gamma"):
                torch.nn.init.constant_(p, 0.3)
    g = torch.Generator().manual_seed(seed + 1)
    imgs = (torch.rand(B, 3, size, size, generator=g) * 2 - 1).cuda()
    words = torch.randn(B, nef, T_, generator=g).cuda()
    sent = torch.randn(B, nef, generator=g)
    sent[2] = sent[5] + 0.05 * torch.randn(nef, generator=g)          # a soft positive for b_global
    lens = torch.randint(2, T_ + 1, (B,), generator=g)
    mask = (torch.arange(T_)[None] >= lens[:, None]).cuda()
    noise = torch.randn(B, 100, generator=g).cuda()
    return G, D, imgs, words, sent.cuda(), mask, noise


def _grads(G, D, inputs, losses, cfg, word_kwargs=None, fused_region_head=False):
    from xmc_gan_b200 import step as S
    imgs, words, sent, mask, noise = inputs
    snap = {}

    def after_d():
        snap["D"] = {n: p.grad.detach().clone() for n, p in D.named_parameters() if p.grad is not None}
    out = S.gd_step(G, D, None, None, imgs, words, sent, mask, noise, cfg=cfg, losses=losses, do_step=False,
                    after_d_backward=after_d, word_kwargs=word_kwargs, fused_region_head=fused_region_head)
    snap["G"] = {n: p.grad.detach().clone() for n, p in G.named_parameters() if p.grad is not None}
    return snap, {k: float(v) for k, v in out.items()}


@pytest.mark.parametrize("b_global,word_precision,tol,fused", [(False, None, 1e-4, False), (True, None, 1e-4, False),
                                                              (False, "bf16", 2e-2, False), (False, "bf16", 2e-2, True)])
def test_swapping_the_loss_ops_leaves_the_step_gradients_unchanged(b_global, word_precision, tol, fused):
    import stock_losses
    from xmc_gan_b200 import step as S
    from xmc_gan_b200 import train_gan as T
    flags = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cuda.matmul.allow_tf32 = False, True, False
    try:
        G, D, *inputs = _setup(B=16, size=64, nch=8, nef=256, T_=7)
        cfg = S.default_step_cfg()
        cfg.TRAIN.ENCODER_LOSS.B_GLOBAL = b_global
        cfg.TRAIN.ENCODER_LOSS.WORD = True
        cfg.TRAIN.SMOOTH.SENT = cfg.TRAIN.SMOOTH.DISC = cfg.TRAIN.SMOOTH.WORD = 1.0
        ref, ref_out = _grads(G, D, inputs, stock_losses, cfg)
        got, got_out = _grads(G, D, inputs, T, cfg, word_kwargs={"precision": word_precision} if word_precision else None,
                              fused_region_head=fused)       # fused: the region head runs in the word loss's prologue (N2)
        for k in ("errD", "errG", "ds_loss", "gs_loss", "disc_loss", "dw_loss", "gw_loss"):
            assert abs(got_out[k] - ref_out[k]) <= tol * max(1.0, abs(ref_out[k])), (k, got_out[k], ref_out[k])
        for which in ("D", "G"):
            assert set(got[which]) == set(ref[which])
            num = sum(float((got[which][n].double() - ref[which][n].double()).pow(2).sum()) for n in ref[which])
            den = sum(float(ref[which][n].double().pow(2).sum()) for n in ref[which])
            assert (num / den) ** 0.5 <= tol, (which, (num / den) ** 0.5)
            for n in ref[which]:                                    # and tensor by tensor, against the update's gradient scale
                scale = max(float(ref[which][n].norm()), 1e-3 * den ** 0.5)
                assert float((got[which][n] - ref[which][n]).norm()) <= 10 * tol * scale, (which, n)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.deterministic, torch.backends.cuda.matmul.allow_tf32 = flags
