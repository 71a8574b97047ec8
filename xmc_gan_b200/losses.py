"""torch.autograd ops of the contrastive-loss path, written against the C-ABI backend (``ops.py``).

Single GPU: every score matrix is square and complete.  With a process ``group`` (global
negatives, one process per GPU): each rank keeps its local batch as the ROWS of every score
matrix, all-gathers only the (small) column operand, owns rows ``rank*B .. (rank+1)*B-1`` of the
global logits, exchanges the per-column statistics (``B_global`` floats x 3) and reduce-scatters
the gradient of the gathered operand back to its owners.  Region features never leave their GPU.
The returned loss is the GLOBAL loss (identical on every rank); gradients are those of the global
loss with respect to the rank's local inputs.

Reference behaviour reproduced (citations into /root/reference/xmc_gan/train_gan.py):
``num_pos`` rule :94-99, column/row log-softmax directions :103-111, ``s0 + s1`` :113.
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.distributed as dist

from . import _lib
from .config import cfg


# ------------------------------------------------------------------------------------------------
# process-group plumbing (NCCL over NVLink on the GPU box; gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------
class Comm:
    def __init__(self, group):
        self.group = group
        self.active = group is not None and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0

    def all_gather_cat(self, t: torch.Tensor) -> torch.Tensor:
        """[n, ...] per rank -> [world*n, ...] in rank order."""
        if not self.active:
            return t
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    def reduce_scatter_sum(self, t_all: torch.Tensor) -> torch.Tensor:
        """[world*n, ...] partial sums on every rank -> this rank's [n, ...] slice of the total."""
        if not self.active:
            return t_all
        t_all = t_all.contiguous()
        n = t_all.shape[0] // self.world
        out = torch.empty((n,) + tuple(t_all.shape[1:]), device=t_all.device, dtype=t_all.dtype)
        dist.reduce_scatter_tensor(out, t_all, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.active:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def combine_col_stats(self, col_stats: torch.Tensor, ops=None) -> torch.Tensor:
        """Per-rank column statistics over local rows -> statistics over all rows.

        Row 0 (log-sum-exp) combines by log-sum-exp across ranks, rows 1-2 (label sums) by sum.
        Traffic: 3 * B_global floats per rank.  The merge is one launch of the backend
        (``xmc_infonce_combine_stats``); the torch expression below serves the CPU checker backend of the tests.
        """
        if not self.active:
            return col_stats
        gathered = self.all_gather_cat(col_stats.unsqueeze(0))          # [world, 3, Bk]
        if ops is not None and hasattr(ops, "combine_col_stats"):
            return ops.combine_col_stats(gathered)
        out = torch.empty_like(col_stats)
        out[0] = torch.logsumexp(gathered[:, 0], dim=0)
        out[1:] = gathered[:, 1:].sum(dim=0)
        return out


def _num_pos(labels, row_count, b_global):
    """Divisor rule of sent_loss / img_loss (train_gan.py:94-99): scalar, or the row-count vector."""
    if not b_global:
        return 1.0, None
    if cfg.TRAIN.SMOOTH.GLOBAL == 0.0:
        return 2.0, None
    if row_count is None:
        row_count = (labels > 0).sum(1).to(torch.float32)
    return 1.0, row_count.contiguous()


def _label_args(labels, comm: Comm, n_rows):
    """-> (dense labels or None, diag offset, row_count or None).

    ``make_labels`` tags identity matrices so the kernels can take NULL + offset instead of reading
    a B x B tensor; anything else is passed dense (fp32, contiguous).
    """
    if labels is None or getattr(labels, "_xmc_identity", False):
        return None, comm.rank * n_rows, None
    rc = getattr(labels, "_xmc_row_count", None)
    if labels.dtype != torch.float32 or not labels.is_contiguous():
        labels = labels.to(torch.float32).contiguous()
    return labels, 0, rc


def _divisors(labels, row_count, b_global, comm: Comm, n_rows, ops):
    """(num_pos scalar, row_div[Bq] or None, col_div[Bk_global] or None).

    The reference divides column j's sum by num_pos[j], the positive count of ROW j (:99,105) —
    kept.  Sharded: row counts of all ranks are all-gathered (B_global floats).
    """
    num_pos, vec = _num_pos(labels, row_count, b_global)
    if vec is None:
        return num_pos, None, None
    return num_pos, vec, comm.all_gather_cat(vec)


# ------------------------------------------------------------------------------------------------
# sentence–image / image–image InfoNCE
# ------------------------------------------------------------------------------------------------
class SimLossFn(torch.autograd.Function):
    """loss = infonce(cosine_scores(a, b) * scale, labels) — train_gan.py:93-115 / 117-139."""

    @staticmethod
    def forward(ctx, a, b, labels, b_global, scale, group, ops):
        comm = Comm(group)
        a_c = a.detach().contiguous()
        b_loc = b.detach().contiguous()
        if a_c.dtype != b_loc.dtype:
            raise TypeError(f"operand dtypes differ: {a_c.dtype} vs {b_loc.dtype}")
        b_all = comm.all_gather_cat(b_loc)
        Bq, Bk = a_c.shape[0], b_all.shape[0]
        lab, diag, rc = _label_args(labels, comm, Bq)
        if lab is not None and tuple(lab.shape) != (Bq, Bk):
            raise ValueError(f"labels must be [{Bq}, {Bk}], got {tuple(lab.shape)}")
        num_pos, row_div, col_div = _divisors(lab, rc, b_global, comm, Bq, ops)
        scores, inv_a, inv_b, row_stats, col_stats = ops.simloss_forward(a_c, b_all, lab, diag, float(scale))
        col_stats = comm.combine_col_stats(col_stats, ops)
        rows_total = Bq * comm.world
        nloc = Bk // comm.world
        loss3 = ops.infonce_loss(row_stats, col_stats, row_div, col_div, num_pos, rows_total, Bk,
                                 comm.rank * nloc, nloc)
        loss3 = comm.all_reduce_sum(loss3)
        ctx.comm, ctx.ops = comm, ops
        ctx.meta = (diag, float(scale), num_pos, rows_total, Bk)
        ctx.save_for_backward(a_c, b_all, scores, inv_a, inv_b, row_stats, col_stats,
                              *(t if t is not None else torch.empty(0) for t in (lab, row_div, col_div)))
        ctx.has = (lab is not None, row_div is not None, col_div is not None)
        return loss3[0]

    @staticmethod
    def backward(ctx, grad_out):
        a, b_all, scores, inv_a, inv_b, row_stats, col_stats, lab, row_div, col_div = ctx.saved_tensors
        lab, row_div, col_div = (t if h else None for t, h in zip((lab, row_div, col_div), ctx.has))
        diag, scale, num_pos, rows_total, cols_total = ctx.meta
        go = grad_out.detach().to(torch.float32).contiguous()
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        da, db_all = ctx.ops.simloss_backward(a, b_all, scores, inv_a, inv_b, lab, diag, scale, row_stats,
                                              col_stats, row_div, col_div, num_pos, rows_total, cols_total,
                                              go, need_a, need_b)
        db = ctx.comm.reduce_scatter_sum(db_all) if need_b else None
        return da, db, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# word–region attention contrastive loss
# ------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def _side_scope(ops, dev):
    """Fork onto the backend's side stream if it has one (the CPU checker backend of the tests has none)."""
    scope = getattr(ops, "side_scope", None)
    if scope is None:
        yield lambda: None
    else:
        with scope(dev) as mark:
            yield mark


def _wait_mark(ops, dev, ev):
    if ev is not None:
        ops.wait_mark(dev, ev)


def _join_side(ops, dev, *tensors):
    if getattr(ops, "join_side", None) is not None:
        ops.join_side(dev, *tensors)



TC_BACKWARD_DIMS = (128, 256)   # D handled by the tcgen05 backward kernel (others: fp32 kernel)


def _ceil_to(x, m):
    return (x + m - 1) // m * m


class WordLossFn(torch.autograd.Function):
    """loss = infonce(rho3 * S_word(regions, words, mask), labels); spec in oracle/word_region.py."""

    @staticmethod
    def forward(ctx, regions, words, mask, labels, b_global, rho1, rho2, rho3, normalize_values,
                precision, group, ops):
        comm = Comm(group)
        reg = regions.detach().flatten(2).contiguous()             # [Bi, D, R]
        w_loc = words.detach().contiguous()                        # [Bc, D, T]
        if reg.dtype != w_loc.dtype:
            raise TypeError(f"operand dtypes differ: {reg.dtype} vs {w_loc.dtype}")
        Bi, D, R = reg.shape
        T = w_loc.shape[2]
        if precision is None:
            precision = "bf16" if reg.dtype == torch.bfloat16 else "fp32"
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32', 'bf16' or None")
        path = _lib.PATH_FP32_SIMT if precision == "fp32" else _lib.PATH_BF16_TCGEN05
        op_dtype = torch.float32 if precision == "fp32" else torch.bfloat16

        w_all = comm.all_gather_cat(w_loc)                         # [Bc_g, D, T]
        Bc = w_all.shape[0]
        if mask is not None:
            m_all = comm.all_gather_cat(mask.detach().to(torch.uint8).contiguous())
        else:
            m_all = None
        Rpad = _ceil_to(R, 16)                                     # zero rows up to the MMA's N granularity
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        use_tc_bwd = path == _lib.PATH_BF16_TCGEN05 and D in TC_BACKWARD_DIMS
        # Padding words never contribute (excluded from the log-sum-exp, zero gradient): on the tcgen05
        # path the kernels visit only the valid word rows, compacted in caption-major order.  Their
        # number stays on the device (cap_ptr[Bc]); nothing here synchronises with the host.
        compact = (m_all is not None and path == _lib.PATH_BF16_TCGEN05 and (use_tc_bwd or not need_grad)
                   and getattr(ops, "supports_compaction", False))
        row_of = cap_ptr = nq_dev = None
        rn_used = not normalize_values
        dev = reg.device
        # The word-side prologue and the zero fill of the backward's accumulators are independent of the
        # region prologue and of the forward kernel: they go to a side stream and are joined below.
        with _side_scope(ops, dev) as mark:
            if compact:
                row_of, cap_ptr = ops.word_rows_compact(m_all)
                nq_dev = cap_ptr[Bc:]
                qn, qnorm = ops.normalize_transpose(w_all, T, op_dtype, row_of=row_of)   # compact rows of [Bc_g*T, D]
            else:
                qn, qnorm = ops.normalize_transpose(w_all, T, op_dtype)                  # [Bc_g, T, D]
            words_ready = mark()
            ctx.bufs = None
            if need_grad and use_tc_bwd and hasattr(ops, "backward_buffers"):
                ctx.bufs = ops.backward_buffers(path, Bc * T, Bi, R, Rpad, D, dev, rn_used)
        kn, rnorm = ops.normalize_transpose(reg, Rpad, op_dtype)   # [Bi, Rpad, D]
        _wait_mark(ops, dev, words_ready)
        qn2 = qn.view(Bc * T, D)
        rn = None if normalize_values else rnorm
        if compact:
            lsum, cnorm, rel, chat = ops.wordregion_forward(path, qn2, kn, rn, R, rho1,
                                                            save_context=need_grad and use_tc_bwd, nq_dev=nq_dev)
            scores = ops.word_scores(rel, m_all, Bc, T, rho2, cap_ptr=cap_ptr)       # [Bi, Bc_g]
        else:
            lsum, cnorm, rel, chat = ops.wordregion_forward(path, qn2, kn, rn, R, rho1,
                                                            save_context=need_grad and use_tc_bwd)
            scores = ops.word_scores(rel, m_all, Bc, T, rho2)      # [Bi, Bc_g]

        lab, diag, rc = _label_args(labels, comm, Bi)
        if lab is not None and tuple(lab.shape) != (Bi, Bc):
            raise ValueError(f"labels must be [{Bi}, {Bc}], got {tuple(lab.shape)}")
        num_pos, row_div, col_div = _divisors(lab, rc, b_global, comm, Bi, ops)
        row_stats, col_stats = ops.infonce_stats(scores, lab, diag, float(rho3))
        col_stats = comm.combine_col_stats(col_stats, ops)
        rows_total = Bi * comm.world
        nloc = Bc // comm.world
        loss3 = ops.infonce_loss(row_stats, col_stats, row_div, col_div, num_pos, rows_total, Bc,
                                 comm.rank * nloc, nloc)
        loss3 = comm.all_reduce_sum(loss3)

        _join_side(ops, dev, qn, qnorm, row_of, cap_ptr, *(ctx.bufs[:4] if ctx.bufs is not None else ()))
        if chat is None:
            ctx.bufs = None
        ctx.comm, ctx.ops = comm, ops
        ctx.meta = (path, R, T, float(rho1), float(rho2), float(rho3), diag, num_pos, rows_total, Bc,
                    tuple(regions.shape), regions.dtype, words.dtype)
        ctx.has = (rn is not None, m_all is not None, lab is not None, row_div is not None, col_div is not None,
                   chat is not None, compact)
        e = torch.empty(0)
        ctx.save_for_backward(qn, qnorm, kn, rnorm, lsum, cnorm, rel, scores, row_stats, col_stats,
                              m_all if m_all is not None else e, lab if lab is not None else e,
                              row_div if row_div is not None else e, col_div if col_div is not None else e,
                              chat if chat is not None else e, row_of if compact else e, cap_ptr if compact else e)
        return loss3[0]

    @staticmethod
    def backward(ctx, grad_out):
        (qn, qnorm, kn, rnorm, lsum, cnorm, rel, scores, row_stats, col_stats,
         m_all, lab, row_div, col_div, chat, row_of, cap_ptr) = ctx.saved_tensors
        has_rn, has_m, has_lab, has_rd, has_cd, has_chat, compact = ctx.has
        m_all = m_all if has_m else None
        lab = lab if has_lab else None
        row_div = row_div if has_rd else None
        col_div = col_div if has_cd else None
        (path, R, T, rho1, rho2, rho3, diag, num_pos, rows_total, Bc, reg_shape, reg_dtype, w_dtype) = ctx.meta
        ops, comm = ctx.ops, ctx.comm
        need_reg, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_reg or need_w):
            return (None,) * 12
        bufs, ctx.bufs = ctx.bufs, None           # single use: the kernels accumulate into them
        go = grad_out.detach().to(torch.float32).contiguous()
        cap = {"cap_ptr": cap_ptr} if compact else {}
        if hasattr(ops, "word_scores_infonce_backward"):     # d loss / d rel in one launch
            grel = ops.word_scores_infonce_backward(rel, m_all, scores, T, rho2, lab, diag, rho3, row_stats, col_stats,
                                                    row_div, col_div, num_pos, rows_total, Bc, go, **cap)
        else:
            dscores = ops.infonce_grad(scores, lab, diag, rho3, row_stats, col_stats, row_div, col_div, num_pos,
                                       rows_total, Bc, go)
            grel = ops.word_scores_backward(rel, m_all, scores, dscores, T, rho2, **cap)
        D = qn.shape[2]
        if compact:
            nq_dev = cap_ptr[Bc:]
            dqn, dkn, drnorm = ops.wordregion_backward(path, qn.view(-1, D), kn, rnorm if has_rn else None, R, rho1,
                                                       lsum, cnorm, rel, grel, chat, nq_dev=nq_dev, bufs=bufs)
        else:
            if path == _lib.PATH_BF16_TCGEN05 and not has_chat:
                # D outside the tcgen05 backward kernel's set: run the fp32 CUDA-core backward kernel on the
                # (bf16-rounded) operands the forward used.  Still libxmcloss, never PyTorch.
                dqn, dkn, drnorm = ops.wordregion_backward(_lib.PATH_FP32_SIMT, qn.view(-1, D).float(), kn.float(),
                                                           rnorm if has_rn else None, R, rho1, lsum, cnorm, rel, grel)
            else:
                dqn, dkn, drnorm = ops.wordregion_backward(path, qn.view(-1, D), kn, rnorm if has_rn else None, R, rho1,
                                                           lsum, cnorm, rel, grel, chat if has_chat else None,
                                                           **({"bufs": bufs} if bufs is not None else {}))
        dreg = dwords = dw_all = None
        dev = kn.device
        if need_w:                                 # the two layout epilogues are independent: words on the side stream
            with _side_scope(ops, dev):
                dw_all = ops.normalize_transpose_backward(qn, qnorm, dqn.view(qn.shape), None, T, torch.float32,
                                                          **({"row_of": row_of} if compact else {}))
        if need_reg:
            dreg = ops.normalize_transpose_backward(kn, rnorm, dkn, drnorm, R, reg_dtype).view(reg_shape)
        if need_w:
            _join_side(ops, dev, dw_all)
            dwords = comm.reduce_scatter_sum(dw_all).to(w_dtype)
        return (dreg, dwords) + (None,) * 10


# ------------------------------------------------------------------------------------------------
# MA-GP reduction (train_gan.py:244-249)
# ------------------------------------------------------------------------------------------------
class GradNormPenaltyFn(torch.autograd.Function):
    """loss = weight * mean_b ||cat(g0[b], g1[b])||_2 ** power, first-order differentiable in g0, g1.

    The gradients fed in carry the graph of ``autograd.grad(..., create_graph=True)``; what this
    function returns from ``backward`` flows on into that graph (the double backward through the
    discriminator is autograd's, as in the reference)."""

    @staticmethod
    def forward(ctx, g0, g1, power, weight, ops):
        a = g0.detach().reshape(g0.shape[0], -1).contiguous()
        b = g1.detach().reshape(g1.shape[0], -1).contiguous()
        if a.shape[0] != b.shape[0]:
            raise ValueError(f"batch sizes differ: {a.shape[0]} vs {b.shape[0]}")
        if a.dtype != b.dtype:
            raise TypeError(f"gradient dtypes differ: {a.dtype} vs {b.dtype}")
        loss, sumsq = ops.gradnorm_penalty_forward(a, b, power, weight)
        ctx.save_for_backward(a, b, sumsq)
        ctx.meta = (float(power), float(weight), tuple(g0.shape), tuple(g1.shape))
        ctx.ops = ops
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        a, b, sumsq = ctx.saved_tensors
        power, weight, s0, s1 = ctx.meta
        go = grad_out.detach().to(torch.float32).contiguous()
        d0, d1 = ctx.ops.gradnorm_penalty_backward(a, b, power, weight, sumsq, go, ctx.needs_input_grad[0],
                                                   ctx.needs_input_grad[1])
        return (d0.view(s0) if d0 is not None else None, d1.view(s1) if d1 is not None else None, None, None, None)
