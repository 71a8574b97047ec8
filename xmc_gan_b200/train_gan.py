"""Drop-in mirror of the reference's loss block (``xmc_gan/train_gan.py:72-139``).

Same module-level names, same positional/keyword parameter names and the same ``cfg`` access
(``cfg.TRAIN.SMOOTH.GLOBAL``), so the reference's call sites (``:213, 218, 265, 278``) work
unchanged after ``from xmc_gan_b200.train_gan import make_labels, cosine_scores, sent_loss,
img_loss, word_loss``.  ``word_loss`` is the loss the reference names (``:222, 269``) but leaves
as ``raise NotImplementedError``.

Everything runs in hand-written sm_100a kernels through ``libxmcloss.so``; non-CUDA tensors raise.
New options are keyword-only and default to the reference's behaviour:

* ``tau``    — temperature of the logits (reference: none, i.e. 1.0);
* ``group``  — a ``torch.distributed`` process group: global negatives across ranks.  NB this is
  NOT the reference's ``b_global`` flag, which means in-batch soft positives (``:72-83``).
"""
from __future__ import annotations

import torch

from . import losses as _L
from .config import cfg  # noqa: F401  (re-exported: callers set cfg.TRAIN.SMOOTH.GLOBAL here)
from .ops import default_ops

__all__ = ["cfg", "make_labels", "cosine_scores", "sent_loss", "img_loss", "word_loss", "contrastive_losses",
           "magp_penalty"]


def make_labels(batch_size, sent_embs, b_global, p=0.6, *, group=None, device=None, _ops=None):
    """Label matrix of the contrastive losses — ``xmc_gan/train_gan.py:72-83``.

    ``b_global=False``: identity ``[B, B]`` (``:74``).  The returned dense tensor is tagged so the
    loss kernels take "identity + offset" instead of reading it.  ``b_global=True``: soft positives
    where ``cos(sent_i, sent_j) > p`` (``:76-82``), computed on the GPU by ``xmc_cosine_scores`` +
    ``xmc_make_labels``; the row counts ``(labels > 0).sum(1)`` come back with it (``:99``).
    With ``group``, returns the rank's rows ``[B, B_global]`` of the global label matrix.
    """
    ops = _ops or default_ops()            # _ops: test hook (CPU checker backend under gloo), never set by users
    comm = _L.as_comm(group)
    if device is None:
        device = sent_embs.device if sent_embs is not None and (sent_embs.is_cuda or _ops) else torch.device("cuda")
    if not b_global:
        B_all = batch_size * comm.world
        labels = torch.zeros(batch_size, B_all, device=device, dtype=torch.float32)
        labels.diagonal(comm.rank * batch_size).fill_(1.0)
        return _L.tag_identity(labels)
    s = sent_embs.detach()
    if s.dtype not in (torch.float32, torch.bfloat16):
        s = s.to(torch.float32)
    s_all = comm.all_gather_cat(s.contiguous())
    sim = ops.cosine_scores(s_all, s_all)
    labels, row_count = ops.make_labels(sim, p, cfg.TRAIN.SMOOTH.GLOBAL)
    col_count = row_count                  # the divisor of column j is the positive count of ROW j (:99, 105)
    if comm.active:
        lo = comm.rank * batch_size
        labels = labels[lo:lo + batch_size].contiguous()
        row_count = row_count[lo:lo + batch_size].contiguous()
    # (labels > 0).sum(1) of the local rows / of every row of the global matrix, valid while the tensor is unedited
    labels._xmc_row_count, labels._xmc_col_count = row_count, col_count
    labels._xmc_counts_of = (labels._version, tuple(labels.shape))
    return labels


def cosine_scores(emb0, emb1, *, _ops=None):
    """``normalize(emb0) @ normalize(emb1).T`` — ``xmc_gan/train_gan.py:85-91``; differentiable like the
    reference's expression (the backward is one kernel: dS @ other side + normalize-backward)."""
    ops = _ops or default_ops()
    if torch.is_grad_enabled() and (emb0.requires_grad or emb1.requires_grad):
        return _L.CosineScoresFn.apply(emb0, emb1, ops)
    return ops.cosine_scores(emb0.detach(), emb1.detach())


def sent_loss(imgs, txts, labels, b_global, *, tau=1.0, group=None, _ops=None):
    """Sentence–image InfoNCE, rows = images, cols = texts — ``xmc_gan/train_gan.py:93-115``."""
    return _L.SimLossFn.apply(imgs, txts, labels, bool(b_global), 1.0 / tau, group, _ops or default_ops())


def img_loss(real_imgs, fake_imgs, labels, b_global, *, tau=1.0, group=None, _ops=None):
    """Real–fake image InfoNCE, rows = real, cols = fake — ``xmc_gan/train_gan.py:117-139``."""
    return _L.SimLossFn.apply(real_imgs, fake_imgs, labels, bool(b_global), 1.0 / tau, group, _ops or default_ops())


def _head_args(imgs, region_head, precision):
    """region_head = (weight, bias) or a module with .weight / .bias (a Conv2d(Cin, D, 1)).  -> (imgs, weight, bias):
    fused (weight given back) on the bf16 tensor-core path, otherwise the projection is applied here with PyTorch's
    convolution and the loss sees ordinary region features."""
    if region_head is None:
        return imgs, None, None
    weight, bias = ((region_head.weight, region_head.bias) if hasattr(region_head, "weight") else region_head)
    fused = (precision == "bf16" or (precision is None and imgs.dtype == torch.bfloat16)) and weight.shape[0] == 256
    if fused:
        return imgs, weight, bias
    w4 = weight.reshape(weight.shape[0], -1, 1, 1)
    x = imgs if imgs.dim() == 4 else imgs.unsqueeze(-1)
    return torch.nn.functional.conv2d(x, w4.to(x.dtype), None if bias is None else bias.to(x.dtype)), None, None


def word_loss(imgs, words, mask, labels, b_global, *, rho1=5.0, rho2=5.0, rho3=10.0,
              normalize_values=False, precision=None, group=None, region_head=None, _ops=None):
    """Word–region attention contrastive loss (name pinned by ``train_gan.py:222, 269``).

    imgs: region features ``[B, D, H, W]`` (or ``[B, D, R]``); words ``[B, D, T]`` and
    mask ``[B, T]`` (True = padding) as produced by the reference's encoders
    (``xmc_gan/model/encoder.py:61,68,140,149``).  Rows = images, cols = captions.
    ``precision``: ``"fp32"`` (rel 1e-4: tcgen05 with every operand carried as a hi + lo bf16 pair for D = 256,
    CUDA-core fp32 kernels otherwise or with ``"fp32-simt"``) or ``"bf16"`` (tcgen05, bf16 operands, fp32
    accumulate, rel 2e-2); default follows the input dtype.
    ``region_head``: ``(weight [D, Cin(,1,1)], bias [D] or None)`` or a ``Conv2d(Cin, D, 1)`` — ``imgs`` is then the
    discriminator's feature map ``[B, Cin, H, W]`` (``xmc_gan/model/df_gan.py:106-132``) and the regions are its 1x1
    projection.  With ``precision="bf16"`` (or bf16 features) and D = 256 the projection, the normalisation and the bf16 cast
    are ONE tensor-core kernel in the loss prologue and the gradients of the map, the weight and the bias
    come out of the loss's backward; otherwise the projection runs as a PyTorch convolution in front of the loss.
    """
    imgs, head_w, head_b = _head_args(imgs, region_head, precision)
    return _L.WordLossFn.apply(imgs, words, mask, labels, bool(b_global), float(rho1), float(rho2), float(rho3),
                               bool(normalize_values), precision, group, _ops or default_ops(), head_w, head_b)


def contrastive_losses(imgs=None, txts=None, real_imgs=None, fake_imgs=None, regions=None, words=None, mask=None,
                       labels=None, b_global=False, *, tau=1.0, rho1=5.0, rho2=5.0, rho3=10.0, normalize_values=False,
                       precision=None, group=None, region_head=None, _ops=None):
    """``(sent_loss(imgs, txts), img_loss(real_imgs, fake_imgs), word_loss(regions, words, mask))`` of one training
    step (``xmc_gan/train_gan.py:218/265, :278, :220-222/267-269``), evaluated together: same numbers as the three
    calls, but with a process ``group`` the collectives are grouped (one all-gather of all column operands, one
    statistics exchange, one reduce-scatter of the gradients) and the similarity losses run on side streams beside
    the word-region kernels.  A pair left ``None`` is skipped and its loss is a constant 0.  ``region_head``: as in
    :func:`word_loss` (``regions`` is then the discriminator's feature map)."""
    head_w = head_b = None
    if regions is not None:
        regions, head_w, head_b = _head_args(regions, region_head, precision)
    return _L.FusedLossesFn.apply(imgs, txts, real_imgs, fake_imgs, regions, words, mask, labels, bool(b_global),
                                  float(tau), float(rho1), float(rho2), float(rho3), bool(normalize_values), precision,
                                  group, _ops or default_ops(), head_w, head_b)


def pooled_features(features, *, out_dtype=None, _ops=None):
    """``F.avg_pool2d(features, kernel_size=H).view(B, -1)`` for a ``[B, C, H, H]`` map — the pooled image embedding that
    feeds ``sent_loss`` (``xmc_gan/model/df_gan.py:165-166``) and both operands of ``img_loss``
    (``xmc_gan/train_gan.py:271-276``) — as one kernel; ``out_dtype=torch.bfloat16`` hands the bf16 similarity path its
    operand without a cast pass."""
    return _L.PooledFeaturesFn.apply(features, out_dtype, _ops or default_ops())


def magp_penalty(grads, *, power=6.0, weight=2.0, _ops=None):
    """Matching-aware gradient penalty reduction, ``xmc_gan/train_gan.py:244-249``.

    ``grads``: the pair returned by ``torch.autograd.grad(out[0], (interpolated, sent_inter),
    create_graph=True, ...)`` (``:237-242``) — image gradients ``[B, 3, H, W]`` and sentence
    gradients ``[B, D]``.  Returns ``weight * mean_b(||cat(g_img[b], g_sent[b])||_2 ** power)``
    (the reference's ``d_loss``: weight 2.0, power 6) as a 0-dim tensor; ``.backward()`` continues
    into the double-backward graph of the discriminator exactly as the reference's expression does.
    One pass over the gradients: no ``cat``, no ``grad ** 2`` temporary.
    """
    g0, g1 = grads
    return _L.GradNormPenaltyFn.apply(g0, g1, float(power), float(weight), _ops or default_ops())
