"""The reference's loss block as stock PyTorch on whatever device the tensors live on (TESTS / BENCH A-B ONLY):
the oracle's restatements (``oracle/ref_losses.py``, pinned on the reference's own functions) under the reference's
call signatures, with ``cfg.TRAIN.SMOOTH.GLOBAL`` read at call time as the reference does.  ``xmc_gan_b200.step.gd_step``
takes this namespace in place of ``xmc_gan_b200.train_gan`` to produce the "before the swap" arm."""
from __future__ import annotations

import torch

import oracle
from xmc_gan_b200.config import cfg


def make_labels(batch_size, sent_embs, b_global, p=0.6):
    sim_in = sent_embs.detach().float()
    labels = oracle.make_labels(batch_size, sim_in.cpu(), b_global, p, cfg.TRAIN.SMOOTH.GLOBAL)
    return labels.to(sent_embs.device)


def sent_loss(imgs, txts, labels, b_global):
    return oracle.sent_loss(imgs, txts, labels, b_global, cfg.TRAIN.SMOOTH.GLOBAL)


def img_loss(real_imgs, fake_imgs, labels, b_global):
    return oracle.img_loss(real_imgs, fake_imgs, labels, b_global, cfg.TRAIN.SMOOTH.GLOBAL)


def word_loss(imgs, words, mask, labels, b_global, rho1=5.0, rho2=5.0, rho3=10.0, normalize_values=False, **_):
    m = mask if mask is None or mask.dtype == torch.bool else mask.bool()
    return _word_loss_on_device(imgs, words, m, labels, b_global, rho1, rho2, rho3, normalize_values)


def _word_loss_on_device(imgs, words, mask, labels, b_global, rho1, rho2, rho3, normalize_values):
    # oracle.word_scores builds its masks on the CPU; restate the few lines device-agnostically (same maths)
    from oracle.ref_losses import infonce_tail, num_pos_of
    from oracle.word_region import _unit_last
    v = imgs.flatten(2).transpose(1, 2)
    e = words.transpose(1, 2)
    vn, en = _unit_last(v), _unit_last(e)
    vals = vn if normalize_values else v
    if mask is None:
        mask = torch.zeros(e.shape[:2], dtype=torch.bool, device=e.device)
    empty = mask.all(dim=1)
    out = []
    for i0 in range(0, v.shape[0], 8):
        s = torch.einsum('ctd,ird->ictr', en, vn[i0:i0 + 8])
        a = torch.softmax(rho1 * s, dim=-1)
        ctx = torch.einsum('ictr,ird->ictd', a, vals[i0:i0 + 8])
        rel = (en.unsqueeze(0) * _unit_last(ctx)).sum(-1)
        z = (rho2 * rel).masked_fill(mask.unsqueeze(0), float('-inf'))
        z = torch.where(empty.view(1, -1, 1), torch.zeros_like(z), z)
        sc = torch.logsumexp(z, dim=-1) / rho2
        out.append(torch.where(empty.view(1, -1), torch.zeros_like(sc), sc))
    S = torch.cat(out, dim=0)
    return infonce_tail(rho3 * S, labels, num_pos_of(labels, b_global, cfg.TRAIN.SMOOTH.GLOBAL))


def magp_penalty(grads, power=6.0, weight=2.0):
    return oracle.magp_penalty(grads[0], grads[1], power, weight)
