"""DF-GAN-shaped generator / discriminator for the train-step harness (BASELINE config 4).  TEST / BENCH
INFRASTRUCTURE, not product: the product is the loss path; these networks are only the G/D that surround it in
``xmc_gan/train_gan.py:187-289``.  ``/root/reference`` does not exist on the GPU box, so the step test and
``bench.py --workload step`` need their own networks; this file restates the reference's architecture
(``xmc_gan/model/df_gan.py``: generator :64-103 + :179-263, discriminator :106-176 + :266-294) with the SAME
parameter names, so a reference ``state_dict`` loads unchanged — ``tests/test_step_harness.py`` does exactly
that in the build container and checks equal outputs against the live reference modules, and pins a golden vector
for everywhere else.

Shapes at IMG.SIZE = 256, NCH = 32 (config 4): the discriminator trunk ends in [B, 512, 4, 4]; its 16x16 stage is
[B, 512, 16, 16] — the 256 regions the word loss attends over (``region_index``), projected to NEF = 256 channels by
the 1x1 ``RegionHead`` (new: the reference has no word loss and therefore no region head).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

# channel multipliers per stage, keyed by image size (df_gan.py:9-62)
_GEN = {256: ([8, 8, 8, 8, 8, 4, 2], [8, 8, 8, 8, 4, 2, 1]),
        128: ([8, 8, 8, 8, 4, 2], [8, 8, 8, 4, 2, 1]),
        64: ([8, 8, 8, 4, 2], [8, 8, 4, 2, 1])}
_DISC = {256: [1, 2, 4, 8, 16, 16, 16], 128: [1, 2, 4, 8, 16, 16], 64: [1, 2, 4, 8, 16]}


def _maybe_sn(m, on):
    return nn.utils.spectral_norm(m) if on else m


class _Film(nn.Module):
    """Sentence-conditioned channel-wise scale and shift (df_gan.py:227-263, there called ``affine``)."""

    def __init__(self, channels, cond_dim):
        super().__init__()
        mlp = lambda: nn.Sequential(OrderedDict(linear1=nn.Linear(cond_dim, 256), relu1=nn.ReLU(inplace=True),
                                                linear2=nn.Linear(256, channels)))
        self.fc_gamma, self.fc_beta = mlp(), mlp()
        nn.init.zeros_(self.fc_gamma.linear2.weight); nn.init.ones_(self.fc_gamma.linear2.bias)
        nn.init.zeros_(self.fc_beta.linear2.weight); nn.init.zeros_(self.fc_beta.linear2.bias)

    def forward(self, x, c):
        return self.fc_gamma(c)[:, :, None, None] * x + self.fc_beta(c)[:, :, None, None]


class _GenStage(nn.Module):
    """Residual generator stage with four conditioned scale/shift layers (df_gan.py:179-224)."""

    def __init__(self, cin, cout, cond_dim, upsample):
        super().__init__()
        self.upsample = upsample
        self.c1, self.c2 = nn.Conv2d(cin, cout, 3, 1, 1), nn.Conv2d(cout, cout, 3, 1, 1)
        self.affine0, self.affine1 = _Film(cin, cond_dim), _Film(cin, cond_dim)
        self.affine2, self.affine3 = _Film(cout, cond_dim), _Film(cout, cond_dim)
        self.gamma = nn.Parameter(torch.zeros(1))
        if cin != cout:
            self.c_sc = nn.Conv2d(cin, cout, 1)

    def forward(self, x, c):
        act = lambda t: F.leaky_relu(t, 0.2)
        h = self.c1(act(self.affine1(act(self.affine0(x, c)), c)))
        h = self.c2(act(self.affine3(act(self.affine2(h, c)), c)))
        out = (self.c_sc(x) if hasattr(self, "c_sc") else x) + self.gamma * h
        return F.interpolate(out, scale_factor=2) if self.upsample else out


class NetG(nn.Module):
    def __init__(self, img_size=256, nch=32, noise_dim=100, text_dim=256, nef=256):
        super().__init__()
        cin, cout = _GEN[img_size]
        self.ngf = nch
        self.proj_noise = nn.Linear(noise_dim, 8 * nch * 16)
        self.proj_sent = nn.Linear(text_dim, nef) if text_dim != nef else nn.Identity()
        n = len(cin)
        self.upblocks = nn.ModuleList(_GenStage(cin[i] * nch, cout[i] * nch, nef, i < n - 1) for i in range(n))
        self.conv_out = nn.Sequential(nn.LeakyReLU(0.2), nn.Conv2d(cout[-1] * nch, 3, 3, 1, 1), nn.Tanh())

    def forward(self, noise, sent_embs, **_):
        h = self.proj_noise(noise).view(noise.shape[0], 8 * self.ngf, 4, 4)
        c = self.proj_sent(sent_embs)
        for blk in self.upblocks:
            h = blk(h, c)
        return self.conv_out(h)


class _DiscStage(nn.Module):
    """Strided residual discriminator stage (df_gan.py:266-294)."""

    def __init__(self, cin, cout, spec_norm):
        super().__init__()
        self.conv_r = nn.Sequential(_maybe_sn(nn.Conv2d(cin, cout, 4, 2, 1, bias=False), spec_norm), nn.LeakyReLU(0.2),
                                    _maybe_sn(nn.Conv2d(cout, cout, 3, 1, 1, bias=False), spec_norm), nn.LeakyReLU(0.2))
        self.conv_s = _maybe_sn(nn.Conv2d(cin, cout, 1), spec_norm)
        self.learned = cin != cout
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        s = self.conv_s(x) if self.learned else x
        return F.avg_pool2d(s, 2) + self.gamma * self.conv_r(x)


class CondHead(nn.Module):
    """Pooled image embedding, projected text/image embedding and the conditional logit (df_gan.py:134-176).
    Only the two projection modes the loss path uses: IMG_MATCH (image 16*ndf -> nef) or SENT_MATCH (text nef -> 16*ndf)."""

    def __init__(self, ndf, nef, img_match=True, spec_norm=False):
        super().__init__()
        self.img_match = img_match
        self.proj_match = _maybe_sn(nn.Linear(ndf * 16, nef) if img_match else nn.Linear(nef, ndf * 16), spec_norm)
        cond = nef if img_match else ndf * 16
        self.joint_conv = nn.Sequential(_maybe_sn(nn.Conv2d(ndf * 16 + cond, ndf * 2, 3, 1, 1, bias=False), spec_norm),
                                        nn.LeakyReLU(0.2),
                                        _maybe_sn(nn.Conv2d(ndf * 2, 1, 4, 1, 0, bias=False), spec_norm))

    def forward(self, x, sent_embs, **_):
        pooled = F.avg_pool2d(x, 4).flatten(1)
        if self.img_match:
            pooled = self.proj_match(pooled)
        else:
            sent_embs = self.proj_match(sent_embs)
        c = sent_embs[:, :, None, None].expand(-1, -1, 4, 4)
        return [self.joint_conv(torch.cat((x, c), 1)), pooled, sent_embs]


class RegionHead(nn.Module):
    """1x1 projection of the discriminator's 16x16 stage to the word embedding width (new)."""

    def __init__(self, cin, nef):
        super().__init__()
        self.proj = nn.Conv2d(cin, nef, 1)

    def forward(self, x):
        return self.proj(x)


class NetD(nn.Module):
    def __init__(self, img_size=256, nch=32, nef=256, img_match=True, spec_norm=False, region_res=16):
        super().__init__()
        mult = _DISC[img_size]
        self.conv_img = _maybe_sn(nn.Conv2d(3, mult[0] * nch, 3, 1, 1), spec_norm)
        self.downblocks = nn.ModuleList(_DiscStage(mult[i - 1] * nch, mult[i] * nch, spec_norm) for i in range(1, len(mult)))
        self.COND_DNET = CondHead(nch, nef, img_match, spec_norm)
        # index of the stage whose output is region_res x region_res; its channels feed the region head
        res, self.region_index = img_size, None
        for i in range(len(self.downblocks)):
            res //= 2
            if res == region_res:
                self.region_index = i
        self.region_head = RegionHead(mult[self.region_index + 1] * nch, nef) if self.region_index is not None else None

    def forward(self, x, with_regions=False, **_):
        h = self.conv_img(x)
        regions = None
        for i, blk in enumerate(self.downblocks):
            h = blk(h)
            if with_regions and i == self.region_index:     # "features": the stage's map, for a head fused into the loss
                regions = h if with_regions == "features" else self.region_head(h)
        return (h, regions) if with_regions else h
