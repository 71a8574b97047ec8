"""Thin tensor-level wrappers over the C ABI (one method per entry point of ``include/xmc_loss.h``).

``CudaOps`` is the only product backend.  It takes CUDA tensors, allocates outputs with the
PyTorch caching allocator (the library never allocates), passes raw device pointers plus the
current stream, and raises on any non-zero status.  There is no CPU path: a non-CUDA tensor is an
error.  The autograd layer (``losses.py``) is written against this interface so that the
multi-rank plumbing can be exercised on CPU/gloo in ``tests/`` with a checker backend injected —
the product never does that.
"""
from __future__ import annotations

import contextlib
import os
import threading

import torch

from . import _lib

_DT = {torch.float32: _lib.XMC_F32, torch.bfloat16: _lib.XMC_BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"xmc_gan_b200 supports float32 and bfloat16 operands, got {t.dtype}") from None


def _p(t):
    return None if t is None else t.data_ptr()


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("xmc_gan_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of the current stream of the current device (one C call: this runs ~25 times per step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class _on:
    """``with _on(t):`` makes t's device current for the launch — a no-op (one C call) when it already is,
    which is the case in a one-process-per-GPU job; torch.cuda.device_of costs ~15 us per use."""
    __slots__ = ("idx", "prev")

    def __init__(self, t):
        self.idx = t.device.index

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.idx is not None and self.idx != self.prev:
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.idx is not None and self.idx != self.prev:
            torch.cuda.set_device(self.prev)
        return False


def _f32c(t):
    return None if t is None else t.to(torch.float32).contiguous()


class _Timed:
    def __init__(self, ops, name):
        self.ops, self.name = ops, name

    def __enter__(self):
        if self.ops.events is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.ops.events is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.ops.events.setdefault(self.name, []).append((self.a, b))
        return False


class CudaOps:
    """Calls into libxmcloss.so.  Stateless; safe from the autograd worker thread."""

    name = "libxmcloss"

    def __init__(self, lib=None):
        self.L = lib if lib is not None else _lib.lib()   # lib: the -DXMC_TEST_HOOKS build (tests / experiments only)
        self.launches = 0      # kernels launched through this backend (bench.py's gpu_launches)
        self.events = None     # name -> [(start, stop)] CUDA events when kernel timing is on
        self.check_errors = bool(os.environ.get("XMC_CHECK_ERRORS"))   # tests: read the kernels' error word back
        self._tls = threading.local()   # fork state of side_scope: forward runs on the caller's thread, backward on autograd's

    @property
    def last_workspace(self):
        """Workspace of the last word-region launch made by THIS thread (None: fp32 path, which has none)."""
        return getattr(self._tls, "last_ws", None)

    @last_workspace.setter
    def last_workspace(self, ws):
        self._tls.last_ws = ws

    def _check(self, rc):
        if rc != 0:
            _lib.check(rc, self.L)

    def enable_timing(self, on=True):
        """Record CUDA events (current stream) around the named hot kernels; see kernel_ms()."""
        self.events = {} if on else None

    def kernel_ms(self):
        """name -> (launches, mean ms) over everything recorded since enable_timing()."""
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / len(v)) for k, v in self.events.items()}

    def _timed(self, name):
        return _Timed(self, name)

    # -- similarity losses ------------------------------------------------------------------
    def cosine_scores(self, a, b, with_norms=False):
        _cuda(a, b)
        a, b = a.contiguous(), b.contiguous()
        if a.dtype != b.dtype:
            raise TypeError(f"operand dtypes differ: {a.dtype} vs {b.dtype}")
        Bq, D = a.shape
        Bk = b.shape[0]
        scores = torch.empty(Bq, Bk, device=a.device, dtype=torch.float32)
        inv_a = torch.empty(Bq, device=a.device, dtype=torch.float32) if with_norms else None
        inv_b = torch.empty(Bk, device=a.device, dtype=torch.float32) if with_norms else None
        with _on(a):
            self._check(self.L.xmc_cosine_scores(_p(a), _p(b), Bq, Bk, D, _dt(a), _p(scores), _p(inv_a), _p(inv_b), _stream()))
        self.launches += 1
        return (scores, inv_a, inv_b) if with_norms else scores

    def cosine_scores_backward(self, a, b, inv_a, inv_b, dscores, need_a, need_b):
        _cuda(a, b, dscores)
        if not (need_a or need_b):
            return None, None
        Bq, D = a.shape
        Bk = b.shape[0]
        da = torch.empty_like(a) if need_a else None
        db = torch.empty_like(b) if need_b else None
        with _on(a):
            self._check(self.L.xmc_cosine_scores_backward(_p(a), _p(b), Bq, Bk, D, _dt(a), _p(inv_a), _p(inv_b), _p(dscores),
                                                          _p(da), _p(db), _stream()))
        self.launches += 1
        return da, db

    def simloss_forward(self, a, b, labels, diag, scale, col_stats=None):
        """col_stats: optional [3, Bk] fp32 view to write the column statistics into (a slice of the exchange packet)."""
        _cuda(a, b, labels)
        Bq, D = a.shape
        Bk = b.shape[0]
        dev = a.device
        scores = torch.empty(Bq, Bk, device=dev, dtype=torch.float32)
        inv_a = torch.empty(Bq, device=dev, dtype=torch.float32)
        inv_b = torch.empty(Bk, device=dev, dtype=torch.float32)
        row_stats = torch.empty(3, Bq, device=dev, dtype=torch.float32)
        if col_stats is None:
            col_stats = torch.empty(3, Bk, device=dev, dtype=torch.float32)
        with _on(a), self._timed("simloss_fwd"):
            self._check(self.L.xmc_simloss_forward(_p(a), _p(b), Bq, Bk, D, _dt(a), _p(labels), diag, scale,
                                                  _p(scores), _p(inv_a), _p(inv_b), _p(row_stats), _p(col_stats),
                                                  _stream()))
        self.launches += 1
        return scores, inv_a, inv_b, row_stats, col_stats

    def simloss_backward(self, a, b, scores, inv_a, inv_b, labels, diag, scale, row_stats, col_stats,
                         row_div, col_div, num_pos, rows_total, cols_total, grad_out, need_a, need_b):
        _cuda(a, b, grad_out)
        Bq, D = a.shape
        Bk = b.shape[0]
        da = torch.empty_like(a) if need_a else None
        db = torch.empty_like(b) if need_b else None
        if not (need_a or need_b):
            return None, None
        nws = self.L.xmc_simloss_workspace_bytes(Bq, Bk, D) if self.use_sim_tc else 0    # > 0: large problem, tensor-core form
        ws = torch.empty(nws, device=a.device, dtype=torch.uint8) if nws else None
        with _on(a), self._timed("simloss_bwd"):
            self._check(self.L.xmc_simloss_backward(
                _p(a), _p(b), Bq, Bk, D, _dt(a), _p(scores), _p(inv_a), _p(inv_b), _p(labels), diag, scale,
                _p(row_stats), _p(col_stats), _p(row_div), _p(col_div), float(num_pos), rows_total, cols_total,
                _p(grad_out), _p(da), _p(db), _p(ws), nws, _stream()))
        self.launches += 1 + (3 if nws else 0)
        return da, db

    # -- InfoNCE tail over a given score matrix -----------------------------------------------
    def infonce_stats(self, scores, labels, diag, scale, col_stats=None):
        _cuda(scores, labels)
        Bq, Bk = scores.shape
        row_stats = torch.empty(3, Bq, device=scores.device, dtype=torch.float32)
        if col_stats is None:
            col_stats = torch.empty(3, Bk, device=scores.device, dtype=torch.float32)
        with _on(scores):
            self._check(self.L.xmc_infonce_stats(_p(scores), Bq, Bk, _p(labels), diag, scale,
                                                _p(row_stats), _p(col_stats), _stream()))
        self.launches += 1
        return row_stats, col_stats

    def combine_col_stats(self, gathered):
        """[world, 3, Bk] per-shard column statistics -> [3, Bk] statistics over all rows (one launch)."""
        _cuda(gathered)
        world, _, Bk = gathered.shape
        out = torch.empty(3, Bk, device=gathered.device, dtype=torch.float32)
        with _on(gathered):
            self._check(self.L.xmc_infonce_combine_stats(_p(gathered), world, Bk, _p(out), _stream()))
        self.launches += 1
        return out

    def infonce_loss(self, row_stats, col_stats, row_div, col_div, num_pos, rows_total, cols_total,
                     col_begin, col_count, error_word=None, out=None):
        """error_word: workspace of the tcgen05 forward that produced the scores (its word 0 is the kernel's
        error flag); a non-zero flag makes the loss NaN on the device."""
        _cuda(row_stats, col_stats)
        if out is None:
            out = torch.empty(3, device=row_stats.device, dtype=torch.float32)
        with _on(row_stats):
            self._check(self.L.xmc_infonce_loss(_p(row_stats), _p(col_stats), row_stats.shape[1], col_stats.shape[1],
                                               _p(row_div), _p(col_div), float(num_pos), rows_total, cols_total,
                                               col_begin, col_count, _p(out), _p(error_word), _stream()))
        self.launches += 1
        return out

    def combine_loss(self, gathered, offset, Bk, col_div, num_pos, cols_total):
        """gathered [world, P] packets, this loss's packet at float `offset` -> (col_stats [3, Bk], loss3 [3])."""
        _cuda(gathered)
        world, stride = gathered.shape
        col_stats = torch.empty(3, Bk, device=gathered.device, dtype=torch.float32)
        out = torch.empty(3, device=gathered.device, dtype=torch.float32)
        with _on(gathered):
            self._check(self.L.xmc_infonce_combine_loss(gathered.data_ptr() + 4 * offset, world, stride, Bk, _p(col_div),
                                                        float(num_pos), cols_total, _p(col_stats), _p(out), _stream()))
        self.launches += 1
        return col_stats, out

    def infonce_grad(self, scores, labels, diag, scale, row_stats, col_stats, row_div, col_div, num_pos,
                     rows_total, cols_total, grad_out):
        _cuda(scores, grad_out)
        Bq, Bk = scores.shape
        ds = torch.empty_like(scores)
        with _on(scores):
            self._check(self.L.xmc_infonce_grad(_p(scores), Bq, Bk, _p(labels), diag, scale, _p(row_stats),
                                               _p(col_stats), _p(row_div), _p(col_div), float(num_pos),
                                               rows_total, cols_total, _p(grad_out), _p(ds), _stream()))
        self.launches += 1
        return ds

    def make_labels(self, sim, p, smooth_global):
        _cuda(sim)
        B = sim.shape[0]
        labels = torch.empty(B, B, device=sim.device, dtype=torch.float32)
        row_count = torch.empty(B, device=sim.device, dtype=torch.float32)
        tmp = torch.empty(B, device=sim.device, dtype=torch.float32)
        with _on(sim):
            self._check(self.L.xmc_make_labels(_p(sim), B, float(p), float(smooth_global), _p(labels),
                                              _p(row_count), _p(tmp), _stream()))
        self.launches += 2
        return labels, row_count

    # -- MA-GP reduction ----------------------------------------------------------------------
    @staticmethod
    def _gp_slices(B, n0):
        """CTAs per row: enough CTAs for every SM (148) a few times over, at least 16 K elements each."""
        return int(max(1, min(64, (4 * 148 + B - 1) // B, max(1, n0 // 16384))))

    def gradnorm_penalty_forward(self, g0, g1, power, weight):
        """g0 [B, n0], g1 [B, n1] (contiguous, same dtype) -> (loss [] fp32, sumsq [B] fp32)."""
        _cuda(g0, g1)
        B, n0, n1 = g0.shape[0], g0.shape[1], g1.shape[1]
        S = self._gp_slices(B, n0)
        partial = torch.empty(B * S, device=g0.device, dtype=torch.float32)
        sumsq = torch.empty(B, device=g0.device, dtype=torch.float32)
        loss = torch.empty((), device=g0.device, dtype=torch.float32)
        with _on(g0), self._timed("gradpen_fwd"):
            self._check(self.L.xmc_gradnorm_penalty_forward(_p(g0), n0, _p(g1), n1, B, _dt(g0), float(power),
                                                           float(weight), S, _p(partial), _p(sumsq), _p(loss), _stream()))
        self.launches += 2
        return loss, sumsq

    def gradnorm_penalty_backward(self, g0, g1, power, weight, sumsq, grad_out, need0, need1):
        _cuda(g0, g1, grad_out)
        B, n0, n1 = g0.shape[0], g0.shape[1], g1.shape[1]
        S = self._gp_slices(B, n0)
        d0 = torch.empty_like(g0) if need0 else None
        d1 = torch.empty_like(g1) if need1 else None
        with _on(g0), self._timed("gradpen_bwd"):
            self._check(self.L.xmc_gradnorm_penalty_backward(_p(g0), n0, _p(g1), n1, B, _dt(g0), float(power),
                                                            float(weight), S, _p(sumsq), _p(grad_out), _p(d0), _p(d1),
                                                            _stream()))
        self.launches += 1
        return d0, d1

    # -- word-region --------------------------------------------------------------------------
    use_sim_tc = True              # similarity-loss backward: hand the library a workspace (tensor-core form for large problems)
    supports_compaction = True     # the tcgen05 kernels visit only the non-padding word rows
    supports_split = True          # fp32 tolerance on the tensor cores (XMC_PATH_FP32_TCGEN05, D = 256)
    use_side_stream = True         # word-side prologue / zero fills / word epilogue beside the main stream

    def word_rows_compact(self, mask_u8):
        """mask [Bc, T] (non-zero = padding) -> row_of [Bc*T] int32 (-1 = dropped), cap_ptr [Bc+1] int32."""
        _cuda(mask_u8)
        Bc, T = mask_u8.shape
        row_of = torch.empty(Bc * T, device=mask_u8.device, dtype=torch.int32)
        cap_ptr = torch.empty(Bc + 1, device=mask_u8.device, dtype=torch.int32)
        with _on(mask_u8):
            self._check(self.L.xmc_word_rows_compact(_p(mask_u8), Bc, T, _p(row_of), _p(cap_ptr), _stream()))
        self.launches += 1
        return row_of, cap_ptr

    def normalize_transpose(self, x, Lpad, out_dtype, row_of=None):
        _cuda(x)
        B, D, L = x.shape
        if row_of is not None:        # compact rows: unwritten rows (dropped words, tail) must read as zeros
            xn = torch.zeros(B, Lpad, D, device=x.device, dtype=out_dtype)
        else:
            xn = torch.empty(B, Lpad, D, device=x.device, dtype=out_dtype)
        norm = torch.empty(B, Lpad, device=x.device, dtype=torch.float32)
        with _on(x):
            self._check(self.L.xmc_normalize_transpose(_p(x), B, D, L, Lpad, _dt(x), _DT[out_dtype], _p(row_of),
                                                      _p(xn), _p(norm), _stream()))
        self.launches += 1
        return xn, norm

    def normalize_transpose_backward(self, xn, norm, dxn, dnorm, L, out_dtype, row_of=None, error_word=None):
        _cuda(xn, dxn)
        B, Lpad, D = xn.shape
        dx = torch.empty(B, D, L, device=xn.device, dtype=out_dtype)
        with _on(xn):
            self._check(self.L.xmc_normalize_transpose_backward(_p(xn), _p(norm), _p(dxn), _p(dnorm), B, D, L, Lpad,
                                                               _dt(xn), _DT[out_dtype], _p(row_of), _p(error_word), _p(dx),
                                                               _stream()))
        self.launches += 1
        return dx

    def normalize_rows(self, x, Lpad, out_dtype):
        """x [B, L, D] (D contiguous: a channels-last map) -> unit rows [B, Lpad, D] + norms [B, Lpad]; no transpose."""
        _cuda(x)
        B, L, D = x.shape
        xn = torch.empty(B, Lpad, D, device=x.device, dtype=out_dtype)
        norm = torch.empty(B, Lpad, device=x.device, dtype=torch.float32)
        with _on(x):
            self._check(self.L.xmc_normalize_rows(_p(x), B, D, L, Lpad, _dt(x), _DT[out_dtype], _p(xn), _p(norm), _stream()))
        self.launches += 1
        return xn, norm

    def normalize_rows_backward(self, xn, norm, dxn, dnorm, L, out_dtype, error_word=None):
        _cuda(xn, dxn)
        B, Lpad, D = xn.shape
        dx = torch.empty(B, L, D, device=xn.device, dtype=out_dtype)
        with _on(xn):
            self._check(self.L.xmc_normalize_rows_backward(_p(xn), _p(norm), _p(dxn), _p(dnorm), B, D, L, Lpad, _dt(xn),
                                                           _DT[out_dtype], _p(error_word), _p(dx), _stream()))
        self.launches += 1
        return dx

    # -- region head fused into the word loss's prologue -------------------------
    def region_head_forward(self, feat, weight, bias, Rpad):
        """feat [B, Cin, R] (fp32 / bf16, NCHW), weight [D, Cin], bias [D] or None -> unit rows kn [B, Rpad, D] bf16 +
        norms [B, Rpad] of y = conv1x1(feat); y itself is never written."""
        _cuda(feat, weight, bias)
        B, Cin, R = feat.shape
        D = weight.shape[0]
        kn = torch.empty(B, Rpad, D, device=feat.device, dtype=torch.bfloat16)
        rnorm = torch.empty(B, Rpad, device=feat.device, dtype=torch.float32)
        with _on(feat), self._timed("region_head_fwd"):
            self._check(self.L.xmc_region_head_forward(_p(feat), _dt(feat), _p(weight), _dt(weight), _p(bias), B, Cin, R, Rpad, D,
                                                       _p(kn), _p(rnorm), _stream()))
        self.launches += 1
        return kn, rnorm

    def region_head_backward(self, feat, weight, dy, need_feat, need_weight, need_bias):
        """dy [B, R, D] (bf16 / fp32) -> (dfeat like feat or None, dweight [D, Cin] fp32 or None, dbias [D] fp32 or None)."""
        _cuda(feat, weight, dy)
        B, Cin, R = feat.shape
        D = weight.shape[0]
        dfeat = dweight = dbias = None
        with _on(feat), self._timed("region_head_bwd"):
            if need_feat:
                dfeat = torch.empty_like(feat)
                self._check(self.L.xmc_region_head_backward_input(_p(weight), _dt(weight), _p(dy), _dt(dy), B, Cin, R, D,
                                                                  _p(dfeat), _dt(dfeat), _stream()))
                self.launches += 1
            if need_weight or need_bias:
                dweight = torch.empty(D, Cin, device=feat.device, dtype=torch.float32)
                dbias = torch.empty(D, device=feat.device, dtype=torch.float32) if need_bias else None
                self._check(self.L.xmc_region_head_backward_weight(_p(feat), _dt(feat), _p(dy), _dt(dy), B, Cin, R, D,
                                                                   _p(dweight), _p(dbias), _stream()))
                self.launches += 1 + (1 if need_bias else 0)
        return dfeat, dweight, dbias

    def avgpool_rows(self, x, out_dtype=None):
        """x [B, C, P] -> [B, C] mean over P (F.avg_pool2d over the whole map + view, df_gan.py:165-166, train_gan.py:271-276)."""
        _cuda(x)
        B, C, P = x.shape
        out = torch.empty(B, C, device=x.device, dtype=out_dtype or x.dtype)
        with _on(x):
            self._check(self.L.xmc_avgpool_rows(_p(x), _dt(x), B, C, P, _p(out), _dt(out), _stream()))
        self.launches += 1
        return out

    def avgpool_rows_backward(self, dout, P, out_dtype):
        _cuda(dout)
        B, C = dout.shape
        dx = torch.empty(B, C, P, device=dout.device, dtype=out_dtype)
        with _on(dout):
            self._check(self.L.xmc_avgpool_rows_backward(_p(dout), _dt(dout), B, C, P, _p(dx), _dt(dx), _stream()))
        self.launches += 1
        return dx

    def _check_error_word(self, ws, what):
        """XMC_CHECK_ERRORS=1 (tests): synchronise and raise if a bounded mbarrier wait of the tcgen05 kernel
        timed out (word 0 of its workspace).  Off by default: the product path never synchronises."""
        if ws is not None and self.check_errors and not torch.cuda.is_current_stream_capturing():
            code = int(ws[:4].view(torch.int32)[0])
            if code != 0:
                raise RuntimeError(f"{what}: pipeline wait timed out inside the kernel (code {code})")

    def _workspace(self, path, NQ, Bi, R, Rpad, D, dev):
        """Caller-owned scratch of the tcgen05 paths; word 0 is the kernel's error flag (0 = ok): the first 64 bytes are
        zeroed (the split path's workspace also holds its operand planes, tens of MB that need no fill)."""
        n = self.L.xmc_wordregion_workspace_bytes(path, NQ, Bi, R, Rpad, D)
        if not n:
            ws = None
        elif n <= (1 << 20):
            ws = torch.zeros(n, device=dev, dtype=torch.uint8)
        else:
            ws = torch.empty(n, device=dev, dtype=torch.uint8)
            ws[:256].zero_()
        self.last_workspace = ws
        return ws, n

    def wordregion_forward(self, path, qn, kn, rnorm, R, rho1, save_context=False, nq_dev=None):
        """-> lsum, cnorm, rel [Bi, NQ] (+ chat [Bi, NQ, D] bf16 on the tcgen05 path when asked).  The kernel's
        workspace (word 0 = its error flag) is ``self.last_workspace`` on the calling thread afterwards.

        nq_dev: device int32 holding the number of valid (compact) rows of qn; rows beyond it are
        neither read nor written by any kernel of the path."""
        _cuda(qn, kn, rnorm)
        NQ, D = qn.shape
        Bi, Rpad, _ = kn.shape
        dev = qn.device
        lsum = torch.empty(Bi, NQ, device=dev, dtype=torch.float32)
        cnorm = torch.empty(Bi, NQ, device=dev, dtype=torch.float32)
        rel = torch.empty(Bi, NQ, device=dev, dtype=torch.float32)
        chat = None
        if save_context and path == _lib.PATH_BF16_TCGEN05:
            chat = torch.empty(Bi, NQ, D, device=dev, dtype=torch.bfloat16)
        elif save_context and path == _lib.PATH_FP32_TCGEN05:      # hi plane, lo plane
            chat = torch.empty(2, Bi, NQ, D, device=dev, dtype=torch.bfloat16)
        ws, n = self._workspace(path, NQ, Bi, R, Rpad, D, dev)
        with _on(qn), self._timed("wordregion_fwd"):
            self._check(self.L.xmc_wordregion_forward(path, _p(qn), _p(kn), _p(rnorm), NQ, Bi, R, Rpad, D, float(rho1),
                                                     _p(lsum), _p(cnorm), _p(rel), _p(chat), _p(nq_dev), _p(ws), n,
                                                     _stream()))
        self.launches += 1
        self._check_error_word(ws, "wordregion_forward")
        return lsum, cnorm, rel, chat

    def backward_buffers(self, path, NQ, Bi, R, Rpad, D, dev, has_rnorm):
        """Zero-filled gradient accumulators + workspace of wordregion_backward, allocated and filled on the
        CURRENT stream.  The word loss calls this inside ``side_scope`` so that the 80 MB fill of dkn runs beside
        the forward kernel instead of in front of the backward kernel.  -> (dqn, dkn, drnorm, ws, ws_bytes)."""
        dqn = torch.zeros(NQ, D, device=dev, dtype=torch.float32)
        dkn = torch.zeros(Bi, Rpad, D, device=dev, dtype=torch.float32)
        drnorm = torch.zeros(Bi, Rpad, device=dev, dtype=torch.float32) if has_rnorm else None
        ws, n = self._workspace(path, NQ, Bi, R, Rpad, D, dev)
        return dqn, dkn, drnorm, ws, n

    # A fork/join pair for independent pieces of one autograd function (capture-safe: the join happens
    # before the function returns, so a CUDA-graph capture never ends with a dangling forked stream).
    @contextlib.contextmanager
    def side_scope(self, dev, idx=0):
        """Body runs on side stream ``idx``, ordered after what the current stream holds so far.  Yields a
        function ``mark()`` -> event recorded on the side stream at that point."""
        if getattr(self, "events", None) is not None or not self.use_side_stream:
            yield lambda: None                 # per-kernel timing (enable_timing): one stream, clean event pairs
            return
        cur = torch.cuda.current_stream(dev)
        side = self._side_stream(dev, idx)
        if cur == side:                        # nested use from a body that already runs there
            yield lambda: None
            return
        side.wait_stream(cur)
        self._forked().add(idx)
        with torch.cuda.stream(side):
            def mark():
                ev = torch.cuda.Event()
                ev.record(side)
                return ev
            yield mark

    def wait_mark(self, dev, ev):
        torch.cuda.current_stream(dev).wait_event(ev)

    def join_side(self, dev, *tensors, idx=0):
        """Current stream waits for everything on side stream ``idx``; ``tensors`` (allocated there) are marked
        as used by the current stream for the caching allocator."""
        forked = self._forked()
        if idx not in forked:
            return
        forked.discard(idx)
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(self._side_stream(dev, idx))
        for t in tensors:
            if t is not None:
                t.record_stream(cur)

    def _forked(self):
        if not hasattr(self._tls, "forked"):
            self._tls.forked = set()
        return self._tls.forked

    def _side_stream(self, dev, idx=0):
        key = (dev.type, dev.index, idx)
        if not hasattr(self, "_sides"):
            self._sides = {}
        if key not in self._sides:
            self._sides[key] = torch.cuda.Stream(device=dev)
        return self._sides[key]

    def wordregion_backward(self, path, qn, kn, rnorm, R, rho1, lsum, cnorm, rel, grel, chat=None, nq_dev=None,
                            bufs=None):
        _cuda(qn, kn, grel)
        NQ, D = qn.shape
        Bi, Rpad, _ = kn.shape
        dev = qn.device
        if bufs is not None:
            dqn, dkn, drnorm, ws, n = bufs         # zero-filled beside the forward (WordLossFn)
            self.last_workspace = ws
        else:
            dqn = torch.zeros(NQ, D, device=dev, dtype=torch.float32)
            dkn = torch.zeros(Bi, Rpad, D, device=dev, dtype=torch.float32)
            drnorm = torch.zeros(Bi, Rpad, device=dev, dtype=torch.float32) if rnorm is not None else None
            ws, n = self._workspace(path, NQ, Bi, R, Rpad, D, dev)
        with _on(qn), self._timed("wordregion_bwd"):
            self._check(self.L.xmc_wordregion_backward(path, _p(qn), _p(kn), _p(rnorm), NQ, Bi, R, Rpad, D, float(rho1),
                                                      _p(lsum), _p(cnorm), _p(rel), _p(chat), _p(grel), _p(dqn), _p(dkn),
                                                      _p(drnorm), _p(nq_dev), _p(ws), n, _stream()))
        self.launches += 1
        self._check_error_word(ws, "wordregion_backward")
        return dqn, dkn, drnorm

    def word_scores(self, rel, mask_u8, Bc, T, rho2, cap_ptr=None):
        _cuda(rel, mask_u8)
        Bi, NQs = rel.shape
        scores = torch.empty(Bi, Bc, device=rel.device, dtype=torch.float32)
        with _on(rel):
            self._check(self.L.xmc_word_scores(_p(rel), _p(mask_u8), _p(cap_ptr), Bi, Bc, T, NQs, float(rho2),
                                              _p(scores), _stream()))
        self.launches += 1
        return scores

    def word_scores_backward(self, rel, mask_u8, scores, dscores, T, rho2, cap_ptr=None):
        _cuda(rel, dscores)
        Bi, Bc = scores.shape
        NQs = rel.shape[1]
        grel = torch.empty_like(rel)
        with _on(rel):
            self._check(self.L.xmc_word_scores_backward(_p(rel), _p(mask_u8), _p(cap_ptr), _p(scores), _p(dscores),
                                                       Bi, Bc, T, NQs, float(rho2), _p(grel), _stream()))
        self.launches += 1
        return grel

    def word_scores_infonce_backward(self, rel, mask_u8, scores, T, rho2, labels, diag, scale, row_stats, col_stats,
                                     row_div, col_div, num_pos, rows_total, cols_total, grad_out, cap_ptr=None):
        """infonce_grad + word_scores_backward in one launch -> grel [Bi, NQs]."""
        _cuda(rel, scores, grad_out)
        Bi, Bc = scores.shape
        NQs = rel.shape[1]
        grel = torch.empty_like(rel)
        with _on(rel):
            self._check(self.L.xmc_word_scores_infonce_backward(
                _p(rel), _p(mask_u8), _p(cap_ptr), _p(scores), Bi, Bc, T, NQs, float(rho2), _p(labels), diag,
                float(scale), _p(row_stats), _p(col_stats), _p(row_div), _p(col_div), float(num_pos), rows_total,
                cols_total, _p(grad_out), _p(grel), _stream()))
        self.launches += 1
        return grel


_default = None


def default_ops() -> CudaOps:
    """The product backend.  Raises when libxmcloss.so is missing or no sm_100 device is current."""
    global _default
    if _default is None:
        if not torch.cuda.is_available():
            raise RuntimeError("xmc_gan_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        ops = CudaOps()
        _lib.check(ops.L.xmc_check_device())
        _default = ops
    return _default
