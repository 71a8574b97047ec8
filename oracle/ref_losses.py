"""CPU restatement of the reference's similarity-matrix contrastive losses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Plain PyTorch on CPU, any
floating dtype (float64 for tight checks).  Each function cites the reference
lines it restates; citations are into ``/root/reference/``.

The reference reads the soft-positive weight from a module-global
``cfg.TRAIN.SMOOTH.GLOBAL`` (``xmc_gan/config/gan.py:41``, default 0.5; every
shipped YAML sets ``0.``).  Here it is the explicit argument ``smooth_global``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_EPS = 1e-12  # F.normalize default, xmc_gan/train_gan.py:88-89


def _unit_rows(x: torch.Tensor) -> torch.Tensor:
    """x / max(||x||_2, eps) per row — ``F.normalize(x, p=2, dim=1)`` (train_gan.py:88-89)."""
    return x / x.norm(dim=1, keepdim=True).clamp_min(_EPS)


def cosine_scores(emb0: torch.Tensor, emb1: torch.Tensor) -> torch.Tensor:
    """[B0,D],[B1,D] -> [B0,B1] cosine matrix (xmc_gan/train_gan.py:85-91)."""
    return _unit_rows(emb0) @ _unit_rows(emb1).t()


def make_labels(batch_size: int, sent_embs: torch.Tensor, b_global: bool,
                p: float = 0.6, smooth_global: float = 0.5) -> torch.Tensor:
    """Identity labels plus optional in-batch soft positives (train_gan.py:72-83).

    Quirks kept on purpose: the result is always float32 (``torch.ones`` default
    dtype, :74); the soft weight ``1/num_pos`` is a ``[B]`` vector that broadcasts
    along COLUMNS (:80-82); ``num_pos = max(count,1)+1`` (:79).
    The reference's hard-coded ``.cuda()`` (:74) is dropped: this is the CPU oracle.
    """
    labels = torch.eye(batch_size, dtype=torch.float32)
    if not b_global:
        return labels
    sim = cosine_scores(sent_embs, sent_embs)
    off_diag = ~torch.eye(batch_size, dtype=torch.bool)
    pos = (sim > p) & off_diag                      # fill_diagonal_(3) & (<3) == drop the diagonal (:77-78)
    count = pos.sum(1).clamp(min=1) + 1              # :79
    if smooth_global != 0.0:
        weight = torch.full((batch_size,), float(smooth_global), dtype=torch.float32)
    else:
        weight = 1.0 / count.to(torch.float32)       # :81
    soft = weight.unsqueeze(0) * pos.to(torch.float32)   # [B] broadcasts over the last dim (:82)
    return (labels + soft).clamp(max=1.0).detach()


def num_pos_of(labels: torch.Tensor, b_global: bool, smooth_global: float):
    """Divisor rule shared by sent_loss / img_loss (train_gan.py:94-99, 118-123)."""
    if not b_global:
        return 1
    if smooth_global == 0.0:
        return 2
    return (labels > 0).sum(1)


def infonce_tail(scores: torch.Tensor, labels: torch.Tensor, num_pos) -> torch.Tensor:
    """Bidirectional label-weighted InfoNCE over a score matrix (train_gan.py:103-113).

    Column direction: log-softmax over dim 0, weighted by labels, summed over
    dim 0, divided elementwise by ``num_pos`` (a scalar or the ROW-count vector,
    indexed here by column position — reproduced, not fixed), mean over columns.
    Row direction: the same over dim 1.
    """
    labels = labels.to(scores.dtype)
    col = -(F.log_softmax(scores, dim=0) * labels).sum(0) / num_pos
    row = -(F.log_softmax(scores, dim=1) * labels).sum(1) / num_pos
    return col.mean() + row.mean()


def sent_loss(imgs, txts, labels, b_global, smooth_global: float = 0.5):
    """Sentence–image InfoNCE, rows = images, cols = texts (train_gan.py:93-115)."""
    return infonce_tail(cosine_scores(imgs, txts), labels,
                        num_pos_of(labels, b_global, smooth_global))


def img_loss(real_imgs, fake_imgs, labels, b_global, smooth_global: float = 0.5):
    """Real–fake image InfoNCE, rows = real, cols = fake (train_gan.py:117-139)."""
    return infonce_tail(cosine_scores(real_imgs, fake_imgs), labels,
                        num_pos_of(labels, b_global, smooth_global))


def magp_penalty(grad_img: torch.Tensor, grad_sent: torch.Tensor, power: float = 6.0, weight: float = 2.0) -> torch.Tensor:
    """MA-GP reduction, ``xmc_gan/train_gan.py:244-249`` (grad0/grad1 views :244-245, cat :246,
    L2 norm :247, mean of the 6th power :248, factor 2 :249)."""
    grad0 = grad_img.reshape(grad_img.size(0), -1)
    grad1 = grad_sent.reshape(grad_sent.size(0), -1)
    grad = torch.cat((grad0, grad1), dim=1)
    grad_l2norm = torch.sqrt(torch.sum(grad ** 2, dim=1))
    return weight * torch.mean(grad_l2norm ** power)
