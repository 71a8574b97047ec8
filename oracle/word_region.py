"""CPU restatement of the word–region attention contrastive loss.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  **PARITY UNPINNED**: the
reference only names this loss and raises ``NotImplementedError``
(``xmc_gan/train_gan.py:220-222, 267-269``); no reference code, test or golden
vector computes it.  The definition below is this repo's frozen specification.
Two of its stages ARE checked against reference code: the attention stage (``attend``)
against ``CondConceptSampler.get_context_embs`` (``xmc_gan/model/concept_gan.py:532-555``) at
``rho1 = 1`` with unit values, and the InfoNCE tail against ``sent_loss``
(``xmc_gan/train_gan.py:93-115``); the composition is what stays unpinned.
It follows

* the XMC-GAN paper's word–region score (attention of each word over the image
  regions, attended context, word/context cosine, log-sum-exp over words);
* input layout of the reference's encoders: ``words [B, D, T]``,
  ``mask [B, T]`` bool with True = padding (``xmc_gan/model/encoder.py:61,68,140,149``);
* cosine through ``F.normalize(p=2)`` (``xmc_gan/train_gan.py:88-89``);
* padding handled by filling with ``-inf`` before the reduction
  (nearest precedent ``xmc_gan/model/concept_gan.py:383-388``);
* the bidirectional label-weighted InfoNCE tail of ``sent_loss``
  (``xmc_gan/train_gan.py:93-115``), rows = images, cols = captions.

For image i (regions v_r, r < R) and caption c (words e_t, t < T):

    s_tr   = cos(e_t, v_r)
    a_tr   = softmax_r(rho1 * s_tr)
    c_t    = sum_r a_tr * v_r            (raw regions as values; or v_r/||v_r||
                                          when ``normalize_values``)
    rel_t  = cos(e_t, c_t)
    S[i,c] = (1/rho2) * log sum_{t not padded} exp(rho2 * rel_t)
    loss   = infonce_tail(rho3 * S, labels, num_pos)

A caption whose words are all padding has no defined score; this spec sets
``S[i,c] = 0`` with zero gradient for it (never NaN).
"""
from __future__ import annotations

import torch

from .ref_losses import _EPS, infonce_tail, num_pos_of


def _unit_last(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=-1, keepdim=True).clamp_min(_EPS)


def attend(en: torch.Tensor, vn: torch.Tensor, vals: torch.Tensor, rho1: float):
    """The attention stage: unit words ``en [Bc, T, D]`` over unit regions ``vn [Bi, R, D]`` with values ``vals [Bi, R, D]``
    -> (cosines ``s [Bi, Bc, T, R]``, attention ``a`` = softmax over the regions of ``rho1 * s``, contexts ``[Bi, Bc, T, D]``).

    PINNED for ``rho1 = 1`` with unit values: the reference's own attention block
    ``CondConceptSampler.get_context_embs`` (``xmc_gan/model/concept_gan.py:532-555``: ``F.normalize`` of queries and keys,
    ``matmul``, ``-inf`` mask fill, ``Softmax`` over the key axis, ``matmul`` with the unit keys) computes exactly this with
    queries = the words of one caption and keys = the regions of one image (``tests/golden/ref_attn_*.npz``,
    ``tests/test_oracle.py::test_attention_stage_matches_reference_*``).  The temperature, the raw-value variant and the
    composition with the steps around it are this repo's specification."""
    s = torch.einsum('ctd,ird->ictr', en, vn)                 # cosines
    a = torch.softmax(rho1 * s, dim=-1)
    ctx = torch.einsum('ictr,ird->ictd', a, vals)
    return s, a, ctx


def word_scores(regions: torch.Tensor, words: torch.Tensor, mask: torch.Tensor | None,
                rho1: float = 5.0, rho2: float = 5.0, normalize_values: bool = False,
                img_block: int = 16) -> torch.Tensor:
    """Word–region score matrix ``S [B_img, B_cap]``.

    regions: ``[Bi, D, R]`` or ``[Bi, D, H, W]``; words: ``[Bc, D, T]``;
    mask: ``[Bc, T]`` bool, True = padding (or None).  Images are processed in
    blocks of ``img_block`` so the ``[Bi, Bc, T, R]`` tensor stays bounded.
    """
    regions = regions.flatten(2)                      # [Bi, D, R]
    v = regions.transpose(1, 2)                       # [Bi, R, D]
    e = words.transpose(1, 2)                         # [Bc, T, D]
    vn, en = _unit_last(v), _unit_last(e)
    vals = vn if normalize_values else v
    Bc, T, _ = e.shape
    if mask is None:
        mask = torch.zeros(Bc, T, dtype=torch.bool)
    empty = mask.all(dim=1)                           # fully padded captions
    out = []
    for i0 in range(0, v.shape[0], img_block):
        vb, vnb, valb = v[i0:i0 + img_block], vn[i0:i0 + img_block], vals[i0:i0 + img_block]
        _, _, ctx = attend(en, vnb, valb, rho1)
        rel = (en.unsqueeze(0) * _unit_last(ctx)).sum(-1)     # [i, c, t]
        z = (rho2 * rel).masked_fill(mask.unsqueeze(0), float('-inf'))
        z = torch.where(empty.view(1, -1, 1), torch.zeros_like(z), z)   # keep LSE finite
        sc = torch.logsumexp(z, dim=-1) / rho2
        sc = torch.where(empty.view(1, -1), torch.zeros_like(sc), sc)
        out.append(sc)
        del vb
    return torch.cat(out, dim=0)


def word_loss(imgs, words, mask, labels, b_global, smooth_global: float = 0.5,
              rho1: float = 5.0, rho2: float = 5.0, rho3: float = 10.0,
              normalize_values: bool = False, img_block: int = 16):
    """Word–region contrastive loss (spec above); name pinned by train_gan.py:222,269."""
    S = word_scores(imgs, words, mask, rho1, rho2, normalize_values, img_block)
    return infonce_tail(rho3 * S, labels, num_pos_of(labels, b_global, smooth_global))
