"""One discriminator + generator update of XMC-GAN with the loss path of this package in it — SURVEY §8(f) N3.

``gd_step`` is the body of the reference's training loop, ``xmc_gan/train_gan.py:187-289``, as a function: the
hinge terms (:195-210, :261), the contrastive losses and their weights (:212-224, :263-284), the label matrix
built once in the D step and reused in the G step (:213 -> :265, :278), the optional matching-aware gradient
penalty (:231-252) and the two optimiser steps (:226-229, :286-289).  It takes the caller's ``netG`` / ``netD``
(anything with the reference's interface: ``netG(noise=, sent_embs=, ...)``, ``netG.proj_sent``, ``netD(imgs)``,
``netD.COND_DNET(features, sent_embs=)``) and a ``losses`` namespace providing ``make_labels, sent_loss, img_loss,
word_loss, magp_penalty`` — ``xmc_gan_b200.train_gan`` by default; the tests pass a stock-PyTorch namespace to
check that swapping the ops leaves every parameter gradient unchanged.

What the reference leaves as ``raise NotImplementedError`` (:220-222, :267-269) is filled in: with
``cfg.TRAIN.ENCODER_LOSS.WORD`` the word loss is evaluated on region features — ``netD(imgs, with_regions=True)``
must then return ``(features, regions [B, NEF, H, W])``.  With ``fused_region_head=True`` the 1x1 region head is not run
by the network: ``netD(imgs, with_regions="features")`` returns the stage's feature map ``[B, Cin, H, W]`` and
``netD.region_head.proj`` (a ``Conv2d(Cin, NEF, 1)``) goes to ``word_loss(region_head=...)``, which projects,
normalises and casts in one tensor-core kernel of the loss prologue.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import train_gan as _T


def default_step_cfg():
    """The knobs ``gd_step`` reads, with the reference's defaults (xmc_gan/config/gan.py:25-43, 65-67)."""
    return SimpleNamespace(
        TRAIN=SimpleNamespace(
            NOISE_DIM=100, N_CRITIC=1, MAGP=True, RMIS_LOSS=False,
            ENCODER_LOSS=SimpleNamespace(B_GLOBAL=False, SENT=True, WORD=False, DISC=True),
            SMOOTH=SimpleNamespace(MISMATCH=0.5, SENT=0.1, DISC=0.1, WORD=0.1)),
        DISC=SimpleNamespace(SEPERATE=False))


def gd_step(netG, netD, optimizerG, optimizerD, imgs, words_embs, sent_embs, mask, noise, cfg=None, losses=None,
            group=None, word_kwargs=None, do_step=True, after_d_backward=None, fused_region_head=False):
    """-> dict of the step's scalars (tensors, no host sync): errD, errG, ds_loss, gs_loss, disc_loss, dw_loss, gw_loss, d_gp.

    ``do_step=False`` skips the optimiser updates and the gradient penalty (gradients stay in ``.grad``);
    ``after_d_backward()`` is called right after ``errD.backward()`` — both serve the parity test, which compares
    the parameter gradients of the two updates between loss implementations."""
    cfg = cfg or default_step_cfg()
    L = losses or _T
    E, S = cfg.TRAIN.ENCODER_LOSS, cfg.TRAIN.SMOOTH
    kw = {} if group is None else {"group": group}
    wkw = dict(word_kwargs or {})
    B = mask.size(0)
    out = {}
    words_embs, sent_embs = words_embs.detach(), sent_embs.detach()                      # :181
    psent = sent_embs if cfg.DISC.SEPERATE else netG.proj_sent(sent_embs)               # :188-191

    # ---- discriminator ----
    want_regions = bool(E.WORD)
    with_regions = True
    if fused_region_head and want_regions:              # N2: the head runs inside the loss prologue
        with_regions = "features"
        wkw["region_head"] = netD.region_head.proj
    real_features = netD(imgs, with_regions=with_regions) if want_regions else netD(imgs)       # :193
    real_features, real_regions = real_features if want_regions else (real_features, None)
    outputs_real = netD.COND_DNET(real_features, sent_embs=psent.detach())               # :194
    errD_real = F.relu(1.0 - outputs_real[0]).mean()                                     # :195
    fake = netG(noise=noise, sent_embs=sent_embs, words_embs=words_embs, mask=mask)      # :199
    fake_features = netD(fake.detach())                                                  # :201
    outputs_fake = netD.COND_DNET(fake_features, sent_embs=psent.detach())
    mis_loss = F.relu(1.0 + outputs_fake[0]).mean()                                      # :204-205
    if cfg.TRAIN.RMIS_LOSS:                                                              # :207-210
        outputs_mis = netD.COND_DNET(real_features[:B - 1], sent_embs=psent[1:B].detach())
        mis_loss = mis_loss + F.relu(1.0 + outputs_mis[0]).mean()
    labels = None
    if E.SENT or E.WORD or E.DISC:
        labels = L.make_labels(B, b_global=E.B_GLOBAL, sent_embs=sent_embs, **kw)        # :212-213
    enc_loss = 0.0
    if E.SENT:
        out["ds_loss"] = L.sent_loss(imgs=outputs_real[1], txts=outputs_real[2], labels=labels, b_global=E.B_GLOBAL, **kw)   # :218
        enc_loss = enc_loss + S.SENT * out["ds_loss"]
    if E.WORD:                                                                           # :220-222 (reference: NotImplementedError)
        out["dw_loss"] = L.word_loss(real_regions, words_embs, mask, labels, E.B_GLOBAL, **wkw, **kw)
        enc_loss = enc_loss + S.WORD * out["dw_loss"]
    errD = errD_real + mis_loss * S.MISMATCH + enc_loss                                  # :224
    netG.zero_grad(); netD.zero_grad()
    errD.backward()
    if after_d_backward is not None:
        after_d_backward()
    if do_step:
        optimizerD.step()
    out["errD"] = errD.detach()

    if cfg.TRAIN.MAGP and do_step:                                                       # :231-252
        interpolated = imgs.detach().requires_grad_()
        sent_inter = psent.detach().requires_grad_()
        o = netD.COND_DNET(netD(interpolated), sent_inter)
        grads = torch.autograd.grad(outputs=o[0], inputs=(interpolated, sent_inter), grad_outputs=torch.ones_like(o[0]),
                                    retain_graph=True, create_graph=True, only_inputs=True)
        d_gp = L.magp_penalty(grads)                                                     # :244-249
        optimizerD.zero_grad(); optimizerG.zero_grad()
        d_gp.backward()
        optimizerD.step()
        out["d_gp"] = d_gp.detach()

    # ---- generator ----
    features = netD(fake, with_regions=with_regions) if want_regions else netD(fake)     # :259
    features, fake_regions = features if want_regions else (features, None)
    outputs = netD.COND_DNET(features, sent_embs=psent)                                  # :260
    errG = -outputs[0].mean()                                                            # :261
    if E.SENT:
        out["gs_loss"] = L.sent_loss(imgs=outputs[1], txts=outputs[2], labels=labels, b_global=E.B_GLOBAL, **kw)     # :265
        errG = errG + S.SENT * out["gs_loss"]
    if E.WORD:                                                                           # :267-269
        out["gw_loss"] = L.word_loss(fake_regions, words_embs, mask, labels, E.B_GLOBAL, **wkw, **kw)
        errG = errG + S.WORD * out["gw_loss"]
    if E.DISC:                                                                           # :270-279
        pool = getattr(L, "pooled_features", None) or (lambda x: F.avg_pool2d(x, kernel_size=4).view(B, -1))
        with torch.no_grad():
            real_pooled = pool(netD(imgs))                                               # :271-273
        fake_pooled = pool(features)                                                     # :275-276
        out["disc_loss"] = L.img_loss(real_imgs=real_pooled, fake_imgs=fake_pooled, labels=labels, b_global=E.B_GLOBAL, **kw)
        errG = errG + S.DISC * out["disc_loss"]
    netG.zero_grad(); netD.zero_grad()
    errG.backward()
    if do_step:
        optimizerG.step()
    out["errG"] = errG.detach()
    return out
