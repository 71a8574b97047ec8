"""The train-step harness (BASELINE config 4 / SURVEY §8f N3), CPU part.

* ``tests/dfgan_harness.py`` restates the reference's DF-GAN generator / discriminator with the same parameter
  names: where ``/root/reference`` exists (the build container) a reference ``state_dict`` is loaded into it and
  the outputs are compared with the live reference modules; everywhere, a committed golden vector
  (``tests/golden/dfgan_harness.npz``, written by this file's ``__main__``) pins them.
* ``xmc_gan_b200.step.gd_step`` with the stock-PyTorch loss namespace runs end to end on CPU and moves parameters.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLDEN = os.path.join(HERE, "golden", "dfgan_harness.npz")
REF_ROOT = "/root/reference"
SIZE, NCH, NEF, NOISE = 64, 4, 24, 10


def _ref_cfg(img_match=True):
    return SimpleNamespace(
        TRAIN=SimpleNamespace(NCH=NCH, NOISE_DIM=NOISE, NEF=NEF), IMG=SimpleNamespace(SIZE=SIZE),
        TEXT=SimpleNamespace(EMBEDDING_DIM=NEF),
        DISC=SimpleNamespace(SPEC_NORM=False, IMG_MATCH=img_match, SENT_MATCH=not img_match, SEPERATE=False))


def _inputs():
    g = torch.Generator().manual_seed(0)
    return (torch.randn(3, NOISE, generator=g), torch.randn(3, NEF, generator=g),
            torch.rand(3, 3, SIZE, SIZE, generator=g) * 2 - 1)


def _harness(seed=0, img_match=True):
    from dfgan_harness import NetD, NetG
    torch.manual_seed(seed)
    G = NetG(SIZE, NCH, NOISE, NEF, NEF)
    D = NetD(SIZE, NCH, NEF, img_match=img_match, spec_norm=False, region_res=16)
    for m in (G, D):                       # the zero-initialised residual gates would hide the residual branches
        for n, p in m.named_parameters():
            if n.endswith("gamma"):
                torch.nn.init.constant_(p, 0.37)
    return G, D


def _outputs(G, D):
    noise, sent, imgs = _inputs()
    with torch.no_grad():
        fake = G(noise=noise, sent_embs=sent)
        feat = D(imgs)
        match, pooled, txt = D.COND_DNET(feat, sent_embs=sent)
    return dict(fake=fake, feat=feat, match=match, pooled=pooled, txt=txt)


@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason="needs the read-only reference (build container only)")
@pytest.mark.parametrize("img_match", [True, False])
def test_harness_equals_the_live_reference_modules(img_match):
    sys.path.insert(0, REF_ROOT)
    try:
        from xmc_gan.model import df_gan as ref
    finally:
        sys.path.remove(REF_ROOT)
    G, D = _harness(img_match=img_match)
    cfg = _ref_cfg(img_match)
    Gr, Dr = ref.NetG(cfg), ref.NetD(cfg)
    Gr.load_state_dict(G.state_dict())                                       # same names, same shapes
    missing = Dr.load_state_dict({k: v for k, v in D.state_dict().items() if not k.startswith("region_head")})
    assert not missing.missing_keys and not missing.unexpected_keys
    noise, sent, imgs = _inputs()
    mine = _outputs(G, D)
    with torch.no_grad():
        fake = Gr(noise=noise, sent_embs=sent)
        feat = Dr(imgs)
        match, pooled, txt = Dr.COND_DNET(feat, sent_embs=sent)
    for k, v in dict(fake=fake, feat=feat, match=match, pooled=pooled, txt=txt).items():
        assert torch.allclose(mine[k], v, atol=1e-6, rtol=1e-5), k


def test_harness_matches_the_golden_vector():
    G, D = _harness()
    g = np.load(GOLDEN)
    for k, v in _outputs(G, D).items():
        assert np.allclose(v.numpy(), g[k], atol=2e-5, rtol=1e-4), k


def test_region_stage_shape():
    _, D = _harness()
    _, _, imgs = _inputs()
    feat, regions = D(imgs, with_regions=True)
    assert feat.shape == (3, 16 * NCH, 4, 4) and regions.shape == (3, NEF, 16, 16)


def test_gd_step_runs_with_stock_losses_on_cpu():
    import stock_losses
    from xmc_gan_b200 import step as S
    G, D = _harness()
    optG = torch.optim.Adam(G.parameters(), 1e-4, betas=(0.0, 0.9))
    optD = torch.optim.Adam(D.parameters(), 4e-4, betas=(0.0, 0.9))
    noise, sent, imgs = _inputs()
    g = torch.Generator().manual_seed(1)
    words = torch.randn(3, NEF, 5, generator=g)
    mask = torch.tensor([[False] * 5, [False, False, True, True, True], [False] * 4 + [True]])
    cfg = S.default_step_cfg()
    cfg.TRAIN.NOISE_DIM = NOISE
    cfg.TRAIN.ENCODER_LOSS.WORD = True
    before = [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())]
    out = S.gd_step(G, D, optG, optD, imgs, words, sent, mask, noise, cfg=cfg, losses=stock_losses)
    for k in ("errD", "errG", "ds_loss", "gs_loss", "disc_loss", "dw_loss", "gw_loss", "d_gp"):
        assert k in out and torch.isfinite(out[k]).all(), k
    moved = sum(int(not torch.equal(a, b.detach())) for a, b in zip(before, list(G.parameters()) + list(D.parameters())))
    assert moved > 10


if __name__ == "__main__":        # regenerate the golden vector (build container; checked against the live reference above)
    G, D = _harness()
    np.savez_compressed(GOLDEN, **{k: v.numpy() for k, v in _outputs(G, D).items()})
    print("wrote", GOLDEN)
