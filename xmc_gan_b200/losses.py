"""torch.autograd ops of the contrastive-loss path, written against the C-ABI backend (``ops.py``).

Single GPU: every score matrix is square and complete.  With a process ``group`` (global
negatives, one process per GPU): each rank keeps its local batch as the ROWS of every score
matrix, all-gathers only the (small) column operand, owns rows ``rank*B .. (rank+1)*B-1`` of the
global logits, exchanges ONE packet per loss (its column statistics over the local rows, ``3 *
B_global`` floats, plus its row-direction partial loss) and reduce-scatters the gradient of the
gathered operand back to its owners.  Region features never leave their GPU.  The returned loss
is the GLOBAL loss (identical on every rank); gradients are those of the global loss with respect
to the rank's local inputs.

``FusedLossesFn`` evaluates the three losses of a training step together: one grouped all-gather
of every column operand (sentence / fake-image embeddings, words, masks), one packet exchange for
the three losses, one grouped reduce-scatter of the three gradients — three collectives per step
instead of four per loss — with the two similarity losses on side streams beside the word-region
kernels.

Reference behaviour reproduced (citations into /root/reference/xmc_gan/train_gan.py):
``num_pos`` rule :94-99, column/row log-softmax directions :103-111, ``s0 + s1`` :113.
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist

from . import _lib
from .config import cfg


# ------------------------------------------------------------------------------------------------
# process-group plumbing (NCCL over NVLink on the GPU box; gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------
class _Done:
    def wait(self):
        return None


class Comm:
    """The collectives of the path over one ``torch.distributed`` group.

    ``gather_begin`` / ``scatter_begin`` start ONE grouped collective for a list of tensors (NCCL:
    ``ncclGroupStart/End`` through c10d's coalescing manager, asynchronous — the returned handle's
    ``wait()`` orders the current stream after it, so independent kernels launched in between overlap
    the transfer); other backends (gloo in the CPU tests) run them one by one.
    """

    def __init__(self, group):
        self.group = group
        self.active = group is not None and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.coalesce = self.active and dist.get_backend(group) == "nccl"

    # primitives (tests may subclass: e.g. host-staged gloo for several ranks on one GPU)
    def _all_gather(self, outs, ins):
        if self.coalesce and len(ins) > 1:
            with dist._coalescing_manager(self.group, async_ops=True) as cm:
                for o, i in zip(outs, ins):
                    dist.all_gather_into_tensor(o, i, group=self.group)
            return cm
        works = [dist.all_gather_into_tensor(o, i, group=self.group, async_op=self.coalesce) for o, i in zip(outs, ins)]
        return _Works(works) if self.coalesce else _Done()

    def _reduce_scatter(self, outs, ins):
        if self.coalesce and len(ins) > 1:
            with dist._coalescing_manager(self.group, async_ops=True) as cm:
                for o, i in zip(outs, ins):
                    dist.reduce_scatter_tensor(o, i, op=dist.ReduceOp.SUM, group=self.group)
            return cm
        works = [dist.reduce_scatter_tensor(o, i, op=dist.ReduceOp.SUM, group=self.group, async_op=self.coalesce)
                 for o, i in zip(outs, ins)]
        return _Works(works) if self.coalesce else _Done()

    def gather_begin(self, tensors):
        """[n_k, ...] per rank -> ([world*n_k, ...] in rank order for every k, handle).  None entries pass through."""
        if not self.active:
            return list(tensors), _Done()
        ins = [None if t is None else t.contiguous() for t in tensors]
        outs = [None if t is None else torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
                for t in ins]
        live = [k for k, t in enumerate(ins) if t is not None]
        work = self._all_gather([outs[k] for k in live], [ins[k] for k in live]) if live else _Done()
        return outs, work

    def scatter_begin(self, tensors):
        """[world*n_k, ...] partial sums on every rank -> (this rank's [n_k, ...] slice of the totals, handle)."""
        if not self.active:
            return list(tensors), _Done()
        ins = [None if t is None else t.contiguous() for t in tensors]
        outs = [None if t is None else torch.empty((t.shape[0] // self.world,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
                for t in ins]
        live = [k for k, t in enumerate(ins) if t is not None]
        work = self._reduce_scatter([outs[k] for k in live], [ins[k] for k in live]) if live else _Done()
        return outs, work

    def all_gather_cat(self, t: torch.Tensor) -> torch.Tensor:
        (out,), work = self.gather_begin([t])
        work.wait()
        return out

    def reduce_scatter_sum(self, t_all: torch.Tensor) -> torch.Tensor:
        (out,), work = self.scatter_begin([t_all])
        work.wait()
        return out


class _Works:
    def __init__(self, works):
        self.works = works

    def wait(self):
        for w in self.works:
            if w is not None:
                w.wait()


def as_comm(group) -> Comm:
    return group if isinstance(group, Comm) else Comm(group)


# ------------------------------------------------------------------------------------------------
# labels and divisors
# ------------------------------------------------------------------------------------------------
def tag_identity(labels: torch.Tensor) -> torch.Tensor:
    """Mark a dense label matrix as "identity with rank offset" (what make_labels returns for
    ``b_global=False``, train_gan.py:74) so the kernels take NULL + offset instead of reading it.  The tag
    records the tensor's version and shape: an in-place edit or a reshaped view falls back to the dense path."""
    labels._xmc_identity = (labels._version, tuple(labels.shape))
    return labels


def _is_identity(labels, shape) -> bool:
    tag = getattr(labels, "_xmc_identity", None)
    return isinstance(tag, tuple) and tag == (labels._version, tuple(labels.shape)) and tuple(labels.shape) == tuple(shape)


def _label_args(labels, comm: Comm, n_rows, n_cols):
    """-> (dense labels or None, diag offset, row_count or None, col_count or None)."""
    if labels is None or _is_identity(labels, (n_rows, n_cols)):
        return None, comm.rank * n_rows, None, None
    if tuple(labels.shape) != (n_rows, n_cols):
        raise ValueError(f"labels must be [{n_rows}, {n_cols}], got {tuple(labels.shape)}")
    rc = getattr(labels, "_xmc_row_count", None)
    cc = getattr(labels, "_xmc_col_count", None)
    tag = getattr(labels, "_xmc_counts_of", None)
    if tag != (labels._version, tuple(labels.shape)):          # edited after make_labels: recount
        rc = cc = None
    if labels.dtype != torch.float32 or not labels.is_contiguous():
        labels = labels.to(torch.float32).contiguous()
    return labels, 0, rc, cc


def _divisors(lab, diag, row_count, col_count, b_global, comm: Comm, n_rows):
    """(num_pos scalar, row_div[Bq] or None, col_div[Bk_global] or None) — train_gan.py:94-99.

    The reference divides column j's sum by num_pos[j], the positive count of ROW j (:99,105) —
    kept.  Sharded: the counts of all rows come with the labels (make_labels computes the global matrix on
    every rank) or, for hand-made labels, from one all-gather of the local counts (B_global floats).
    """
    if not b_global:
        return 1.0, None, None
    if cfg.TRAIN.SMOOTH.GLOBAL == 0.0:
        return 2.0, None, None
    if lab is None:                                            # identity labels: one positive per row
        return 1.0, None, None
    if row_count is None:
        row_count = (lab > 0).sum(1).to(torch.float32)
    row_count = row_count.contiguous()
    if col_count is None:
        col_count = comm.all_gather_cat(row_count)
    return 1.0, row_count, col_count.contiguous()


PACKET_PAD = 4     # floats after the 3*Bk column statistics: [row-direction partial loss, 0, same, pad]


def _packet_floats(Bk):
    return 3 * Bk + PACKET_PAD


class _Tail:
    """InfoNCE-tail state of one loss between its local forward, the exchange and its backward."""
    __slots__ = ("lab", "diag", "num_pos", "row_div", "col_div", "row_stats", "col_stats", "rows_total", "Bk", "scale")


def _tail_local(ops, comm, tl: _Tail, packet, error_word=None):
    """Row-direction partial loss of the local rows -> the packet's tail (sharded runs only)."""
    if packet is not None:
        ops.infonce_loss(tl.row_stats, tl.col_stats, tl.row_div, tl.col_div, tl.num_pos, tl.rows_total, tl.Bk,
                         0, 0, error_word=error_word, out=packet[3 * tl.Bk:3 * tl.Bk + 3])


def _tail_finish(ops, comm, tl: _Tail, gathered, offset, error_word=None):
    """-> global loss (0-dim).  One GPU: straight from the statistics.  Sharded: merge the gathered packets
    (column log-sum-exps by log-sum-exp, label sums by sum, row partials by sum); every rank evaluates all
    columns, so no all-reduce follows."""
    if gathered is None:
        loss3 = ops.infonce_loss(tl.row_stats, tl.col_stats, tl.row_div, tl.col_div, tl.num_pos, tl.rows_total, tl.Bk,
                                 0, tl.Bk, error_word=error_word)
    else:
        tl.col_stats, loss3 = ops.combine_loss(gathered, offset, tl.Bk, tl.col_div, tl.num_pos, tl.Bk)
    return loss3[0]


# ------------------------------------------------------------------------------------------------
# sentence–image / image–image InfoNCE
# ------------------------------------------------------------------------------------------------
class _Sim:
    __slots__ = ("a", "b_all", "scores", "inv_a", "inv_b", "tail")


def _sim_local(ops, comm, a_c, b_all, labels, b_global, scale, packet):
    Bq, Bk = a_c.shape[0], b_all.shape[0]
    if a_c.dtype != b_all.dtype:
        raise TypeError(f"operand dtypes differ: {a_c.dtype} vs {b_all.dtype}")
    tl = _Tail()
    tl.lab, tl.diag, rc, cc = _label_args(labels, comm, Bq, Bk)
    tl.num_pos, tl.row_div, tl.col_div = _divisors(tl.lab, tl.diag, rc, cc, b_global, comm, Bq)
    tl.rows_total, tl.Bk, tl.scale = Bq * comm.world, Bk, float(scale)
    col_out = packet[:3 * Bk].view(3, Bk) if packet is not None else None
    st = _Sim()
    st.a, st.b_all, st.tail = a_c, b_all, tl
    st.scores, st.inv_a, st.inv_b, tl.row_stats, tl.col_stats = ops.simloss_forward(a_c, b_all, tl.lab, tl.diag, tl.scale,
                                                                                   col_stats=col_out)
    _tail_local(ops, comm, tl, packet)
    return st


def _sim_backward(ops, st: _Sim, go, need_a, need_b):
    tl = st.tail
    return ops.simloss_backward(st.a, st.b_all, st.scores, st.inv_a, st.inv_b, tl.lab, tl.diag, tl.scale, tl.row_stats,
                                tl.col_stats, tl.row_div, tl.col_div, tl.num_pos, tl.rows_total, tl.Bk, go, need_a, need_b)


def _exchange(comm: Comm, packet):
    """[P] floats per rank -> [world, P] (None on one GPU)."""
    if not comm.active:
        return None
    return comm.all_gather_cat(packet.view(1, -1))


class SimLossFn(torch.autograd.Function):
    """loss = infonce(cosine_scores(a, b) * scale, labels) — train_gan.py:93-115 / 117-139."""

    @staticmethod
    def forward(ctx, a, b, labels, b_global, scale, group, ops):
        comm = as_comm(group)
        a_c = a.detach().contiguous()
        b_all = comm.all_gather_cat(b.detach().contiguous())
        packet = torch.empty(_packet_floats(b_all.shape[0]), device=a_c.device, dtype=torch.float32) if comm.active else None
        st = _sim_local(ops, comm, a_c, b_all, labels, b_global, scale, packet)
        loss = _tail_finish(ops, comm, st.tail, _exchange(comm, packet), 0)
        ctx.comm, ctx.ops, ctx.st = comm, ops, st
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        st = ctx.st
        go = grad_out.detach().to(torch.float32).contiguous()
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        da, db_all = _sim_backward(ctx.ops, st, go, need_a, need_b)
        db = ctx.comm.reduce_scatter_sum(db_all) if need_b else None
        return da, db, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# word–region attention contrastive loss
# ------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def _side_scope(ops, dev, idx=0):
    """Fork onto a side stream of the backend if it has one (the CPU checker backend of the tests has none)."""
    scope = getattr(ops, "side_scope", None)
    if scope is None:
        yield lambda: None
    else:
        with scope(dev, idx) as mark:
            yield mark


def _wait_mark(ops, dev, ev):
    if ev is not None:
        ops.wait_mark(dev, ev)


def _join_side(ops, dev, *tensors, idx=0):
    if getattr(ops, "join_side", None) is not None:
        ops.join_side(dev, *tensors, idx=idx)


TC_BACKWARD_DIMS = (128, 256)   # D handled by the bf16 tcgen05 backward kernel (others: fp32 CUDA-core kernel)
SPLIT_DIMS = (256,)             # D handled by the fp32-tolerance tcgen05 path (split-bf16 operands); others: CUDA cores
# Saved attended contexts of the tcgen05 path: [images, word rows of the global batch, D] bf16 — O(B * B_global * T * D),
# 0.6 GB at 256 x 256 x 18 x 256, 4.8 GB per rank at 8 x 256, 17 GB at 1024 x 1024 x 32.  Refuse to grow silently past this many bytes.
MAX_CONTEXT_BYTES = 32 << 30


def _ceil_to(x, m):
    return (x + m - 1) // m * m


class _Word:
    __slots__ = ("path", "R", "T", "rho1", "rho2", "rho3", "qn", "qnorm", "kn", "rnorm", "has_rn", "lsum", "cnorm", "rel",
                 "scores", "m_all", "chat", "row_of", "cap_ptr", "compact", "bufs", "tail", "reg_shape", "reg_dtype", "w_dtype",
                 "Bc", "fwd_ws", "rows_layout", "head", "dhead")


def _rows_view(regions):
    """[B, D, H, W] in channels-last memory (what a 1x1 region head run as a GEMM writes, or any
    ``memory_format=torch.channels_last`` producer) -> its [B, H*W, D] row view, else None.  SURVEY §8f N2: such a
    producer already emits the kernels' layout, so the word loss skips both transposing layout kernels."""
    if (regions.dim() == 4 and regions.shape[2] * regions.shape[3] > 1 and not regions.is_contiguous()
            and regions.is_contiguous(memory_format=torch.channels_last) and regions.shape[1] % 4 == 0):
        B, D, H, W = regions.shape
        return regions.permute(0, 2, 3, 1).reshape(B, H * W, D)              # a view: D is the contiguous axis
    return None


def _head_operands(regions, head):
    """(feat [B, Cin, R] contiguous, weight [D, Cin], bias [D] fp32 or None) of a fused region head."""
    weight, bias = head
    feat = regions.detach().flatten(2).contiguous()
    w2 = weight.detach().reshape(weight.shape[0], -1).to(feat.dtype).contiguous()    # one dtype per product: the cp.async path
    if w2.shape[1] != feat.shape[1]:
        raise ValueError(f"region_head weight {tuple(weight.shape)} does not take {feat.shape[1]} input channels")
    b1 = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return feat, w2, b1


def _word_prepare_regions(ops, regions, precision, head=None):
    """Region prologue (independent of the gathered words: runs while the all-gather is in flight).
    -> (shape [Bi, D, R], input dtype, precision, operand dtype, Rpad, unit rows kn, norms, rows_layout, head operands).

    head = (weight, bias): ``regions`` is the discriminator's feature map [B, Cin, H, W] and the region features are its
    1x1 projection — one tcgen05 kernel writes their unit rows and norms directly; bf16 tolerance."""
    if head is not None:
        if precision not in (None, "bf16"):
            raise ValueError("a fused region head runs on the bf16 tensor-core path (precision=None or 'bf16')")
        feat, w2, b1 = _head_operands(regions, head)
        Bi, _, R = feat.shape
        D = w2.shape[0]
        Rpad = _ceil_to(R, 16)
        kn, rnorm = ops.region_head_forward(feat, w2, b1, Rpad)              # [Bi, Rpad, D] bf16, [Bi, Rpad]
        return (Bi, D, R), None, "bf16", torch.bfloat16, Rpad, kn, rnorm, True, (feat, w2)
    regions = regions.detach()
    rows = _rows_view(regions) if hasattr(ops, "normalize_rows") else None
    if precision is None:
        precision = "bf16" if regions.dtype == torch.bfloat16 else "fp32"
    if precision not in ("fp32", "bf16", "fp32-simt"):
        raise ValueError("precision must be 'fp32', 'bf16', 'fp32-simt' or None")
    op_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    if rows is not None:
        Bi, R, D = rows.shape
        Rpad = _ceil_to(R, 16)
        kn, rnorm = ops.normalize_rows(rows, Rpad, op_dtype)                  # [Bi, Rpad, D], no transpose
    else:
        reg = regions.flatten(2).contiguous()                                # [Bi, D, R]
        Bi, D, R = reg.shape
        Rpad = _ceil_to(R, 16)                                               # zero rows up to the MMA's N granularity
        kn, rnorm = ops.normalize_transpose(reg, Rpad, op_dtype)             # [Bi, Rpad, D]
    return (Bi, D, R), regions.dtype, precision, op_dtype, Rpad, kn, rnorm, rows is not None, None


def _word_local(ops, comm, prep, regions, w_all, m_all, labels, b_global, rho1, rho2, rho3, normalize_values, need_grad,
                packet):
    (Bi, D, R), reg_dtype, precision, op_dtype, Rpad, kn, rnorm, rows_layout, head_ops = prep
    if reg_dtype is not None and reg_dtype != w_all.dtype:
        raise TypeError(f"operand dtypes differ: {reg_dtype} vs {w_all.dtype}")
    Bc, _, T = w_all.shape
    if precision == "bf16":
        path = _lib.PATH_BF16_TCGEN05
    elif precision == "fp32" and D in SPLIT_DIMS and getattr(ops, "supports_split", False):
        path = _lib.PATH_FP32_TCGEN05          # fp32 tolerance on the tensor cores: hi + lo bf16 operands, three MMAs per product
    else:
        path = _lib.PATH_FP32_SIMT             # "fp32-simt", or a width the split path does not take
    use_tc_bwd = (path == _lib.PATH_BF16_TCGEN05 and D in TC_BACKWARD_DIMS) or path == _lib.PATH_FP32_TCGEN05
    # Padding words never contribute (excluded from the log-sum-exp, zero gradient): the kernels visit
    # only the valid word rows, compacted in caption-major order.  Their number stays on the device
    # (cap_ptr[Bc]); nothing here synchronises with the host.  (Not with the tcgen05 forward followed by
    # the fp32 backward — D outside the tensor-core backward's set —: that pair shares dense rows.)
    compact = (m_all is not None and (path == _lib.PATH_FP32_SIMT or use_tc_bwd or not need_grad)
               and getattr(ops, "supports_compaction", False))
    save_ctx = need_grad and use_tc_bwd
    ctx_bytes = Bi * Bc * T * D * (4 if path == _lib.PATH_FP32_TCGEN05 else 2)
    if save_ctx and ctx_bytes > MAX_CONTEXT_BYTES:
        raise RuntimeError(f"word_loss would save {ctx_bytes / 2**30:.1f} GiB of attended contexts "
                           f"([{Bi}, {Bc * T}, {D}]) for its backward; raise xmc_gan_b200.losses.MAX_CONTEXT_BYTES "
                           "or split the batch")
    st = _Word()
    st.row_of = st.cap_ptr = nq_dev = None
    rn_used = not normalize_values
    dev = kn.device
    # The word-side prologue and the zero fill of the backward's accumulators are independent of the
    # forward kernel's other operand: they go to a side stream and are joined below.
    with _side_scope(ops, dev) as mark:
        if compact:
            st.row_of, st.cap_ptr = ops.word_rows_compact(m_all)
            nq_dev = st.cap_ptr[Bc:]
            qn, qnorm = ops.normalize_transpose(w_all, T, op_dtype, row_of=st.row_of)   # compact rows of [Bc_g*T, D]
        else:
            qn, qnorm = ops.normalize_transpose(w_all, T, op_dtype)                     # [Bc_g, T, D]
        words_ready = mark()
        st.bufs = None
        if save_ctx and hasattr(ops, "backward_buffers"):
            st.bufs = ops.backward_buffers(path, Bc * T, Bi, R, Rpad, D, dev, rn_used)
    _wait_mark(ops, dev, words_ready)
    rn = None if normalize_values else rnorm
    st.lsum, st.cnorm, st.rel, st.chat = ops.wordregion_forward(path, qn.view(Bc * T, D), kn, rn, R, rho1,
                                                                save_context=save_ctx, nq_dev=nq_dev)
    st.fwd_ws = getattr(ops, "last_workspace", None)          # word 0 = the kernel's error flag (tcgen05 path)
    st.scores = ops.word_scores(st.rel, m_all, Bc, T, rho2, cap_ptr=st.cap_ptr)           # [Bi, Bc_g]

    tl = _Tail()
    tl.lab, tl.diag, rc, cc = _label_args(labels, comm, Bi, Bc)
    tl.num_pos, tl.row_div, tl.col_div = _divisors(tl.lab, tl.diag, rc, cc, b_global, comm, Bi)
    tl.rows_total, tl.Bk, tl.scale = Bi * comm.world, Bc, float(rho3)
    col_out = packet[:3 * Bc].view(3, Bc) if packet is not None else None
    tl.row_stats, tl.col_stats = ops.infonce_stats(st.scores, tl.lab, tl.diag, tl.scale, col_stats=col_out)
    _tail_local(ops, comm, tl, packet, error_word=st.fwd_ws)

    _join_side(ops, dev, qn, qnorm, st.row_of, st.cap_ptr, *(st.bufs[:4] if st.bufs is not None else ()))
    if st.chat is None:
        st.bufs = None
    st.path, st.R, st.T, st.rho1, st.rho2, st.rho3 = path, R, T, float(rho1), float(rho2), float(rho3)
    st.qn, st.qnorm, st.kn, st.rnorm, st.has_rn = qn, qnorm, kn, rnorm, rn is not None
    st.m_all, st.compact, st.tail, st.Bc = m_all, compact, tl, Bc
    st.reg_shape, st.reg_dtype, st.w_dtype, st.rows_layout = tuple(regions.shape), regions.dtype, w_all.dtype, rows_layout
    st.head, st.dhead = head_ops, None
    return st


def _word_backward(ops, st: _Word, go, need_reg, need_w, need_head=(False, False)):
    """-> (d regions in the caller's layout or None, d gathered words [Bc_g, D, T] fp32 or None).  The word-side layout
    epilogue runs on the side stream; the caller joins it (``_join_side``) before using the second result.  With a fused
    region head ``st.dhead`` = (d weight [D, Cin] fp32, d bias [D] fp32) afterwards (``need_head``)."""
    tl = st.tail
    bufs, st.bufs = st.bufs, None             # single use: the kernels accumulate into them
    cap = {"cap_ptr": st.cap_ptr} if st.compact else {}
    T, Bc = st.T, st.Bc
    if hasattr(ops, "word_scores_infonce_backward"):     # d loss / d rel in one launch
        grel = ops.word_scores_infonce_backward(st.rel, st.m_all, st.scores, T, st.rho2, tl.lab, tl.diag, st.rho3, tl.row_stats,
                                                tl.col_stats, tl.row_div, tl.col_div, tl.num_pos, tl.rows_total, Bc, go, **cap)
    else:
        dscores = ops.infonce_grad(st.scores, tl.lab, tl.diag, st.rho3, tl.row_stats, tl.col_stats, tl.row_div, tl.col_div,
                                   tl.num_pos, tl.rows_total, Bc, go)
        grel = ops.word_scores_backward(st.rel, st.m_all, st.scores, dscores, T, st.rho2, **cap)
    qn, kn = st.qn, st.kn
    D = qn.shape[2]
    rn = st.rnorm if st.has_rn else None
    if st.compact:
        dqn, dkn, drnorm = ops.wordregion_backward(st.path, qn.view(-1, D), kn, rn, st.R, st.rho1, st.lsum, st.cnorm, st.rel,
                                                   grel, st.chat, nq_dev=st.cap_ptr[Bc:], bufs=bufs)
    elif st.path == _lib.PATH_BF16_TCGEN05 and st.chat is None:       # (the split path always has its contexts: D = 256 only)
        # D outside the tcgen05 backward kernel's set: run the fp32 CUDA-core backward kernel on the
        # (bf16-rounded) operands the forward used.  Still libxmcloss, never PyTorch.
        dqn, dkn, drnorm = ops.wordregion_backward(_lib.PATH_FP32_SIMT, qn.view(-1, D).float(), kn.float(), rn, st.R,
                                                   st.rho1, st.lsum, st.cnorm, st.rel, grel)
    else:
        dqn, dkn, drnorm = ops.wordregion_backward(st.path, qn.view(-1, D), kn, rn, st.R, st.rho1, st.lsum, st.cnorm, st.rel,
                                                   grel, st.chat, **({"bufs": bufs} if bufs is not None else {}))
    ws = getattr(ops, "last_workspace", None)                 # word 0 = the backward kernel's error flag
    dreg = dw_all = None
    dev = kn.device
    if need_w:                                 # the two layout epilogues are independent: words on the side stream
        with _side_scope(ops, dev):
            dw_all = ops.normalize_transpose_backward(qn, st.qnorm, dqn.view(qn.shape), None, T, torch.float32, error_word=ws,
                                                      **({"row_of": st.row_of} if st.compact else {}))
    if st.head is not None:                    # fused region head: d y rows (bf16) -> the head's two backward products
        if need_reg or any(need_head):
            feat, w2 = st.head
            dy = ops.normalize_rows_backward(kn, st.rnorm, dkn, drnorm, st.R, feat.dtype, error_word=ws)    # [B, R, D]
            dfeat, dwgt, dbias = ops.region_head_backward(feat, w2, dy, need_reg, need_head[0], need_head[1])
            dreg = dfeat.view(st.reg_shape) if dfeat is not None else None
            st.dhead = (dwgt, dbias)
    elif need_reg and st.rows_layout:          # gradient in the producer's own (channels-last) layout, no transpose
        B_, D_, H_, W_ = st.reg_shape
        dreg = ops.normalize_rows_backward(kn, st.rnorm, dkn, drnorm, st.R, st.reg_dtype, error_word=ws)
        dreg = dreg.view(B_, H_, W_, D_).permute(0, 3, 1, 2)
    elif need_reg:
        dreg = ops.normalize_transpose_backward(kn, st.rnorm, dkn, drnorm, st.R, st.reg_dtype, error_word=ws).view(st.reg_shape)
    return dreg, dw_all


def _head_grads(st, meta, need_head):
    """(d weight, d bias) of a fused region head in the parameters' own shapes and dtypes."""
    if meta is None or st is None or st.dhead is None:
        return (None, None)
    (wshape, wdtype, bdtype), (dwgt, dbias) = meta, st.dhead
    return (dwgt.view(wshape).to(wdtype) if need_head[0] and dwgt is not None else None,
            dbias.to(bdtype) if need_head[1] and dbias is not None else None)


def _mask_u8(mask):
    return None if mask is None else mask.detach().to(torch.uint8).contiguous()


class WordLossFn(torch.autograd.Function):
    """loss = infonce(rho3 * S_word(regions, words, mask), labels); spec in oracle/word_region.py."""

    @staticmethod
    def forward(ctx, regions, words, mask, labels, b_global, rho1, rho2, rho3, normalize_values,
                precision, group, ops, head_w=None, head_b=None):
        comm = as_comm(group)
        (w_all, m_all), work = comm.gather_begin([words.detach(), _mask_u8(mask)])      # in flight during the region prologue
        prep = _word_prepare_regions(ops, regions, precision, None if head_w is None else (head_w, head_b))
        work.wait()
        need_grad = any(ctx.needs_input_grad[i] for i in (0, 1, 12, 13))
        ctx.head_meta = None if head_w is None else (head_w.shape, head_w.dtype, None if head_b is None else head_b.dtype)
        packet = (torch.empty(_packet_floats(w_all.shape[0]), device=regions.device, dtype=torch.float32)
                  if comm.active else None)
        st = _word_local(ops, comm, prep, regions, w_all.contiguous(), m_all, labels, b_global, rho1, rho2, rho3,
                         normalize_values, need_grad, packet)
        loss = _tail_finish(ops, comm, st.tail, _exchange(comm, packet), 0, error_word=st.fwd_ws)
        ctx.comm, ctx.ops, ctx.st = comm, ops, st
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        need_reg, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_head = (ctx.needs_input_grad[12], ctx.needs_input_grad[13])
        if not (need_reg or need_w or any(need_head)):
            return (None,) * 14
        st = ctx.st
        ops, comm = ctx.ops, ctx.comm
        go = grad_out.detach().to(torch.float32).contiguous()
        dreg, dw_all = _word_backward(ops, st, go, need_reg, need_w, need_head)
        dwords = None
        if need_w:
            _join_side(ops, dreg.device if dreg is not None else dw_all.device, dw_all)
            dwords = comm.reduce_scatter_sum(dw_all).to(st.w_dtype)
        return (dreg, dwords) + (None,) * 10 + _head_grads(st, ctx.head_meta, need_head)


# ------------------------------------------------------------------------------------------------
# the three losses of one training step, evaluated together
# ------------------------------------------------------------------------------------------------
class FusedLossesFn(torch.autograd.Function):
    """(sent_loss, img_loss, word_loss) of one step — train_gan.py:218/265, :278 and :220-222/267-269 — with the
    collectives of the sharded path grouped: one all-gather of [txts, fake_imgs, words, mask], one packet exchange,
    one reduce-scatter of [d txts, d fake_imgs, d words].  A loss whose operands are None is skipped (its output is
    a zero that carries no gradient).  The similarity losses run on side streams beside the word-region kernels."""

    @staticmethod
    def forward(ctx, imgs, txts, real_imgs, fake_imgs, regions, words, mask, labels, b_global,
                tau, rho1, rho2, rho3, normalize_values, precision, group, ops, head_w=None, head_b=None):
        comm = as_comm(group)
        has_sent = imgs is not None and txts is not None
        has_img = real_imgs is not None and fake_imgs is not None
        has_word = regions is not None and words is not None
        det = lambda t: None if t is None else t.detach().contiguous()
        (txt_all, fake_all, w_all, m_all), work = comm.gather_begin(
            [det(txts) if has_sent else None, det(fake_imgs) if has_img else None,
             det(words) if has_word else None, _mask_u8(mask) if has_word else None])
        head = None if head_w is None else (head_w, head_b)
        prep = _word_prepare_regions(ops, regions, precision, head) if has_word else None  # overlaps the gather
        ctx.head_meta = None if head_w is None else (head_w.shape, head_w.dtype, None if head_b is None else head_b.dtype)
        work.wait()
        dev = (imgs if has_sent else real_imgs if has_img else regions).device
        Bk = [t.shape[0] if h else 0 for t, h in ((txt_all, has_sent), (fake_all, has_img), (w_all, has_word))]
        offs = [0, _packet_floats(Bk[0]) if has_sent else 0]
        offs.append(offs[1] + (_packet_floats(Bk[1]) if has_img else 0))
        total = offs[2] + (_packet_floats(Bk[2]) if has_word else 0)
        packet = torch.empty(total, device=dev, dtype=torch.float32) if comm.active else None
        sl = (lambda k: packet[offs[k]:offs[k] + _packet_floats(Bk[k])]) if comm.active else (lambda k: None)
        sims = [None, None]
        marks = [None, None]
        for k, (h, a, b_all) in enumerate(((has_sent, imgs, txt_all), (has_img, real_imgs, fake_all))):
            if h:
                with _side_scope(ops, dev, 1 + k) as mark:
                    sims[k] = _sim_local(ops, comm, det(a), b_all, labels, b_global, 1.0 / tau, sl(k))
                    marks[k] = mark
        wst = None
        if has_word:
            need_grad = any(ctx.needs_input_grad[i] for i in (4, 5, 17, 18))
            wst = _word_local(ops, comm, prep, regions, w_all.contiguous(), m_all, labels, b_global, rho1, rho2, rho3,
                              normalize_values, need_grad, sl(2))
        for k in range(2):
            if sims[k] is not None:
                _join_side(ops, dev, sims[k].scores, sims[k].inv_a, sims[k].inv_b, sims[k].tail.row_stats,
                           sims[k].tail.col_stats, idx=1 + k)
        gathered = _exchange(comm, packet)
        zero = None
        out = []
        for k, st in enumerate((sims[0], sims[1], wst)):
            if st is None:
                zero = torch.zeros((), device=dev, dtype=torch.float32) if zero is None else zero
                out.append(zero.clone())
            else:
                out.append(_tail_finish(ops, comm, st.tail, gathered, offs[k], error_word=wst.fwd_ws if k == 2 else None))
        ctx.comm, ctx.ops, ctx.state, ctx.dev = comm, ops, (sims[0], sims[1], wst), dev
        ctx.dtypes = (txts.dtype if has_sent else None, fake_imgs.dtype if has_img else None)
        ctx.set_materialize_grads(False)
        return tuple(out)

    @staticmethod
    def backward(ctx, g_sent, g_img, g_word):
        ops, comm, dev = ctx.ops, ctx.comm, ctx.dev
        s_sent, s_img, wst = ctx.state
        need = ctx.needs_input_grad
        f32 = lambda g: g.detach().to(torch.float32).contiguous()
        d_imgs = d_txt_all = d_real = d_fake_all = dreg = dw_all = None
        # similarity losses on the side streams, the word-region chain on the caller's stream
        if s_sent is not None and g_sent is not None and (need[0] or need[1]):
            with _side_scope(ops, dev, 1):
                d_imgs, d_txt_all = _sim_backward(ops, s_sent, f32(g_sent), need[0], need[1])
        if s_img is not None and g_img is not None and (need[2] or need[3]):
            with _side_scope(ops, dev, 2):
                d_real, d_fake_all = _sim_backward(ops, s_img, f32(g_img), need[2], need[3])
        need_head = (need[17], need[18])
        if wst is not None and g_word is not None and (need[4] or need[5] or any(need_head)):
            dreg_pending = _word_backward(ops, wst, f32(g_word), need[4], need[5], need_head)
            dreg, dw_all = dreg_pending
            if dw_all is not None:
                _join_side(ops, dev, dw_all)
        _join_side(ops, dev, d_imgs, d_txt_all, idx=1)
        _join_side(ops, dev, d_real, d_fake_all, idx=2)
        (d_txt, d_fake, d_words), work = comm.scatter_begin([d_txt_all, d_fake_all, dw_all])
        work.wait()
        if d_words is not None:
            d_words = d_words.to(wst.w_dtype)
        return (d_imgs, d_txt, d_real, d_fake, dreg, d_words) + (None,) * 11 + _head_grads(wst, ctx.head_meta, need_head)


# ------------------------------------------------------------------------------------------------
# cosine_scores as an ordinary differentiable function (train_gan.py:85-91)
# ------------------------------------------------------------------------------------------------
class CosineScoresFn(torch.autograd.Function):
    """scores = normalize(emb0) @ normalize(emb1).T with the gradient PyTorch's autograd gives the reference's
    expression: d emb = normalize-backward(dS @ emb1_hat) (and the transpose for emb1)."""

    @staticmethod
    def forward(ctx, a, b, ops):
        a_c, b_c = a.detach().contiguous(), b.detach().contiguous()
        scores, inv_a, inv_b = ops.cosine_scores(a_c, b_c, with_norms=True)
        ctx.save_for_backward(a_c, b_c, inv_a, inv_b)
        ctx.ops = ops
        return scores

    @staticmethod
    def backward(ctx, g):
        a, b, inv_a, inv_b = ctx.saved_tensors
        da, db = ctx.ops.cosine_scores_backward(a, b, inv_a, inv_b, g.detach().to(torch.float32).contiguous(),
                                                ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return da, db, None


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


# ------------------------------------------------------------------------------------------------
# pooled image embedding (df_gan.py:165-166, train_gan.py:271-276)
# ------------------------------------------------------------------------------------------------
class PooledFeaturesFn(torch.autograd.Function):
    """``F.avg_pool2d(x, kernel_size=H).view(B, -1)`` of a ``[B, C, H, W]`` map with H == W == the window, as one kernel
    (optionally emitting bf16 for the bf16 similarity path)."""

    @staticmethod
    def forward(ctx, x, out_dtype, ops):
        B, C = x.shape[:2]
        xc = x.detach().contiguous().view(B, C, -1)
        ctx.ops, ctx.shape, ctx.dtype = ops, tuple(x.shape), x.dtype
        return ops.avgpool_rows(xc, out_dtype)

    @staticmethod
    def backward(ctx, g):
        P = 1
        for d in ctx.shape[2:]:
            P *= d
        dx = ctx.ops.avgpool_rows_backward(g.detach().contiguous(), P, ctx.dtype)
        return dx.view(ctx.shape), None, None
