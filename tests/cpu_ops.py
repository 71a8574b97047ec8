"""CPU checker backend with the interface of ``xmc_gan_b200.ops.CudaOps`` (TESTS ONLY).

It lets the multi-rank host logic of ``xmc_gan_b200.losses`` (all-gather of column operands,
column-statistics exchange, rank-offset identity labels, reduce-scatter of gradients) run under
``gloo`` on CPU with world_size 2.  Every block op is a plain PyTorch restatement of what the CUDA
kernel behind the same-named C-ABI entry point computes.  The product never imports this file.
"""
from __future__ import annotations

import torch

EPS = 1e-12


def _labels(labels, Bq, Bk, diag, dtype):
    if labels is not None:
        return labels.to(dtype)
    L = torch.zeros(Bq, Bk, dtype=dtype)
    idx = torch.arange(Bq)
    ok = (idx + diag) < Bk
    L[idx[ok], idx[ok] + diag] = 1
    return L


def _stats(Z, L):
    row = torch.stack([torch.logsumexp(Z, 1), L.sum(1), (L * Z).sum(1)])
    col = torch.stack([torch.logsumexp(Z, 0), L.sum(0), (L * Z).sum(0)])
    return row, col


def _dz(Z, L, row_stats, col_stats, row_div, col_div, num_pos, rows_total, cols_total):
    nr = (row_div if row_div is not None else torch.full((Z.shape[0],), float(num_pos))).to(Z.dtype).view(-1, 1)
    nc = (col_div if col_div is not None else torch.full((Z.shape[1],), float(num_pos))).to(Z.dtype).view(1, -1)
    pr = torch.exp(Z - row_stats[0].view(-1, 1))
    pc = torch.exp(Z - col_stats[0].view(1, -1))
    return (pc * col_stats[1].view(1, -1) - L) / (nc * cols_total) + (pr * row_stats[1].view(-1, 1) - L) / (nr * rows_total)


class CpuOps:
    name = "cpu-checker"

    def __init__(self):
        self.launches = 0
        self.compactions = 0

    # -- similarity losses --
    def cosine_scores(self, a, b, with_norms=False):
        inv_a = 1 / a.norm(dim=1).clamp_min(EPS)
        inv_b = 1 / b.norm(dim=1).clamp_min(EPS)
        s = (a * inv_a[:, None]) @ (b * inv_b[:, None]).t()
        return (s, inv_a, inv_b) if with_norms else s

    def cosine_scores_backward(self, a, b, inv_a, inv_b, dscores, need_a, need_b):
        ah, bh = a * inv_a[:, None], b * inv_b[:, None]
        dS = dscores.to(a.dtype)

        def nb(g, xh, inv):
            return (g - xh * (g * xh).sum(1, keepdim=True)) * inv[:, None]
        return (nb(dS @ bh, ah, inv_a) if need_a else None), (nb(dS.t() @ ah, bh, inv_b) if need_b else None)

    def simloss_forward(self, a, b, labels, diag, scale, col_stats=None):
        inv_a = 1 / a.norm(dim=1).clamp_min(EPS)
        inv_b = 1 / b.norm(dim=1).clamp_min(EPS)
        scores = (a * inv_a[:, None]) @ (b * inv_b[:, None]).t()
        L = _labels(labels, a.shape[0], b.shape[0], diag, scores.dtype)
        row, col = _stats(scale * scores, L)
        if col_stats is not None:            # the exchange packet is fp32, as on the GPU
            col_stats.copy_(col)
        return scores, inv_a, inv_b, row, col

    def simloss_backward(self, a, b, scores, inv_a, inv_b, labels, diag, scale, row_stats, col_stats,
                         row_div, col_div, num_pos, rows_total, cols_total, grad_out, need_a, need_b):
        L = _labels(labels, a.shape[0], b.shape[0], diag, scores.dtype)
        dS = grad_out * scale * _dz(scale * scores, L, row_stats, col_stats, row_div, col_div, num_pos, rows_total, cols_total)
        ah, bh = a * inv_a[:, None], b * inv_b[:, None]

        def nb(g, xh, inv):
            return (g - xh * (g * xh).sum(1, keepdim=True)) * inv[:, None]
        da = nb(dS @ bh, ah, inv_a) if need_a else None
        db = nb(dS.t() @ ah, bh, inv_b) if need_b else None
        return da, db

    # -- tail --
    def infonce_stats(self, scores, labels, diag, scale, col_stats=None):
        L = _labels(labels, scores.shape[0], scores.shape[1], diag, scores.dtype)
        row, col = _stats(scale * scores, L)
        if col_stats is not None:
            col_stats.copy_(col)
        return row, col

    def infonce_loss(self, row_stats, col_stats, row_div, col_div, num_pos, rows_total, cols_total, col_begin, col_count,
                     error_word=None, out=None):
        nr = row_div if row_div is not None else float(num_pos)
        s1 = ((row_stats[0] * row_stats[1] - row_stats[2]) / nr).sum() / rows_total
        sl = slice(col_begin, col_begin + col_count)
        nc = col_div[sl] if col_div is not None else float(num_pos)
        s0 = ((col_stats[0][sl] * col_stats[1][sl] - col_stats[2][sl]) / nc).sum() / cols_total
        res = torch.stack([s0 + s1, s0, s1])
        if out is not None:
            out.copy_(res)
            return out
        return res

    def combine_loss(self, gathered, offset, Bk, col_div, num_pos, cols_total):
        """xmc_infonce_combine_loss: merge the ranks' packets, evaluate the global loss."""
        g = gathered[:, offset:offset + 3 * Bk + 1].double()
        cs = g[:, :3 * Bk].reshape(-1, 3, Bk)
        col = torch.stack([torch.logsumexp(cs[:, 0], 0), cs[:, 1].sum(0), cs[:, 2].sum(0)])
        nc = col_div.double() if col_div is not None else float(num_pos)
        s0 = ((col[0] * col[1] - col[2]) / nc).sum() / cols_total
        s1 = g[:, 3 * Bk].sum()
        return col, torch.stack([s0 + s1, s0, s1])

    def infonce_grad(self, scores, labels, diag, scale, row_stats, col_stats, row_div, col_div, num_pos,
                     rows_total, cols_total, grad_out):
        L = _labels(labels, scores.shape[0], scores.shape[1], diag, scores.dtype)
        return grad_out * scale * _dz(scale * scores, L, row_stats, col_stats, row_div, col_div, num_pos, rows_total, cols_total)

    def make_labels(self, sim, p, smooth_global):
        B = sim.shape[0]
        pos = (sim > p) & ~torch.eye(B, dtype=torch.bool)
        count = pos.sum(1).clamp(min=1) + 1
        w = torch.full((B,), float(smooth_global)) if smooth_global != 0 else 1.0 / count.float()
        labels = (torch.eye(B) + w.view(1, -1) * pos.float()).clamp(max=1)
        return labels, (labels > 0).sum(1).float()

    # -- word-region --
    supports_compaction = True

    def word_rows_compact(self, mask_u8):
        """xmc_word_rows_compact: caption-major exclusive scan of the valid words."""
        self.compactions += 1
        Bc, T = mask_u8.shape
        valid = (mask_u8 == 0).flatten()
        excl = torch.cumsum(valid.int(), 0) - valid.int()
        row_of = torch.where(valid, excl, torch.full_like(excl, -1)).to(torch.int32)
        per_cap = valid.view(Bc, T).sum(1)
        cap_ptr = torch.zeros(Bc + 1, dtype=torch.int32)
        cap_ptr[1:] = torch.cumsum(per_cap, 0)
        return row_of, cap_ptr

    def normalize_transpose(self, x, Lpad, out_dtype, row_of=None):
        B, D, L = x.shape
        norm = x.norm(dim=1).clamp_min(EPS)                      # [B, L]
        xn = torch.zeros(B, Lpad, D, dtype=x.dtype)
        n = torch.zeros(B, Lpad, dtype=x.dtype)
        if row_of is not None:                                   # valid rows packed to the front of [B*L, D]
            assert Lpad == L
            keep = row_of >= 0
            dst = row_of[keep].long()
            xn.view(B * L, D)[dst] = (x / norm[:, None, :]).transpose(1, 2).reshape(B * L, D)[keep]
            n.view(B * L)[dst] = norm.reshape(B * L)[keep]
            return xn, n
        xn[:, :L] = (x / norm[:, None, :]).transpose(1, 2)
        n[:, :L] = norm
        return xn, n

    def normalize_transpose_backward(self, xn, norm, dxn, dnorm, L, out_dtype, row_of=None, error_word=None):
        if row_of is not None:                                   # dropped (padding) words get a zero gradient
            B, Lp, D = xn.shape
            assert Lp == L and dnorm is None
            keep = row_of >= 0
            src = row_of.clamp_min(0).long()
            xh, g, n = xn.view(B * L, D)[src], dxn.reshape(B * L, D)[src], norm.view(B * L)[src]
            dx = (g - xh * (g * xh).sum(-1, keepdim=True)) / n.clamp_min(EPS)[:, None]
            dx = torch.where(keep[:, None], dx, torch.zeros_like(dx))
            return dx.view(B, L, D).transpose(1, 2).contiguous()
        xh, g, n = xn[:, :L], dxn[:, :L], norm[:, :L]
        dx = (g - xh * (g * xh).sum(-1, keepdim=True)) / n[..., None]
        if dnorm is not None:
            dx = dx + dnorm[:, :L, None] * xh
        return dx.transpose(1, 2).contiguous()

    def normalize_rows(self, x, Lpad, out_dtype):
        return self.normalize_transpose(x.transpose(1, 2), Lpad, out_dtype)

    def normalize_rows_backward(self, xn, norm, dxn, dnorm, L, out_dtype, error_word=None):
        return self.normalize_transpose_backward(xn, norm, dxn, dnorm, L, out_dtype).transpose(1, 2).contiguous()

    @staticmethod
    def _wr(qn, kn, rnorm, R, rho1):
        s = torch.einsum('qd,ird->iqr', qn, kn[:, :R])
        p = torch.exp(rho1 * (s - 1))
        lsum = p.sum(-1)
        pw = p * (rnorm[:, None, :R] if rnorm is not None else 1)
        ctx = torch.einsum('iqr,ird->iqd', pw, kn[:, :R]) / lsum[..., None]
        cnorm = ctx.norm(dim=-1)
        rel = (qn[None] * ctx).sum(-1) / cnorm.clamp_min(EPS)
        return lsum, cnorm, rel

    def wordregion_forward(self, path, qn, kn, rnorm, R, rho1, save_context=False, nq_dev=None):
        lsum, cnorm, rel = self._wr(qn, kn, rnorm, R, rho1)
        chat = None
        if nq_dev is not None:                # rows beyond the device-side count are never written by the kernel:
            dead = torch.arange(qn.shape[0]) >= int(nq_dev[0])            # poison them so a consumer would notice
            lsum, cnorm, rel = (torch.where(dead[None], torch.full_like(t, float('nan')), t) for t in (lsum, cnorm, rel))
        if save_context:                      # stands for the saved attended contexts; only its presence matters here
            chat = torch.zeros(1)
        return lsum, cnorm, rel, chat

    def wordregion_backward(self, path, qn, kn, rnorm, R, rho1, lsum, cnorm, rel, grel, chat=None, nq_dev=None,
                            bufs=None):
        if nq_dev is not None:                # the kernel visits only the first nq rows
            nq = int(nq_dev[0])
            dq, dk, drn = self.wordregion_backward(path, qn[:nq], kn, rnorm, R, rho1, None, None, None, grel[:, :nq])
            dq_full = torch.zeros_like(qn)
            dq_full[:nq] = dq
            return dq_full, dk, drn
        with torch.enable_grad():            # we are inside an autograd backward: grad mode is off
            q = qn.detach().clone().requires_grad_()
            k = kn.detach().clone().requires_grad_()
            rn = rnorm.detach().clone().requires_grad_() if rnorm is not None else None
            _, _, r = self._wr(q, k, rn, R, rho1)
            (r * grel).sum().backward()
        return q.grad, k.grad, (rn.grad if rn is not None else None)

    @staticmethod
    def _dense(rel, Bc, T, cap_ptr, fill):
        """compact [Bi, NQs] -> caption-padded [Bi, Bc, T] plus the validity mask of its slots."""
        Bi = rel.shape[0]
        cnt = (cap_ptr[1:] - cap_ptr[:-1]).long()
        slot = torch.arange(T)[None, :] < cnt[:, None]                         # [Bc, T]
        src = (cap_ptr[:-1].long()[:, None] + torch.arange(T)[None, :]).clamp_max(rel.shape[1] - 1)
        z = torch.where(slot[None], rel[:, src], torch.full((), fill, dtype=rel.dtype))
        return z, slot, src

    def word_scores(self, rel, mask_u8, Bc, T, rho2, cap_ptr=None):
        if cap_ptr is not None:
            z, slot, _ = self._dense(rel, Bc, T, cap_ptr, float('-inf'))
            empty = ~slot.any(1)
            z = torch.where(empty.view(1, -1, 1), torch.zeros_like(z), rho2 * z)
            return torch.where(empty.view(1, -1), torch.zeros_like(z[..., 0]), torch.logsumexp(z, -1) / rho2)
        z = rho2 * rel.view(rel.shape[0], Bc, T)
        if mask_u8 is not None:
            m = mask_u8.bool()
            empty = m.all(1)
            z = z.masked_fill(m[None], float('-inf'))
            z = torch.where(empty.view(1, -1, 1), torch.zeros_like(z), z)
            return torch.where(empty.view(1, -1), torch.zeros_like(z[..., 0]), torch.logsumexp(z, -1) / rho2)
        return torch.logsumexp(z, -1) / rho2

    def word_scores_backward(self, rel, mask_u8, scores, dscores, T, rho2, cap_ptr=None):
        Bi, Bc = scores.shape
        if cap_ptr is not None:
            z, slot, src = self._dense(rel, Bc, T, cap_ptr, 0.0)
            w = torch.where(slot[None], torch.exp(rho2 * (z - scores[..., None])), torch.zeros_like(z))
            g = dscores[..., None] * w                                          # [Bi, Bc, T]
            grel = torch.zeros_like(rel)
            grel[:, src[slot]] = g[:, slot]
            return grel
        w = torch.exp(rho2 * (rel.view(Bi, Bc, T) - scores[..., None]))
        if mask_u8 is not None:
            w = w.masked_fill(mask_u8.bool()[None], 0.0)
        return (dscores[..., None] * w).reshape(Bi, Bc * T)
