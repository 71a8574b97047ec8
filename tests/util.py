"""Shared helpers of the parity tests (tests only)."""
import numpy as np
import torch

TOL_FP32 = 1e-4     # north_star: fp32 rel <= 1e-4
TOL_BF16 = 2e-2     # north_star: bf16-input / fp32-accumulate rel <= 2e-2


def t(x, dtype=None, device=None):
    out = torch.from_numpy(np.asarray(x))
    if dtype is not None:
        out = out.to(dtype)
    return out.to(device) if device is not None else out


def nerr(x, ref):
    """Norm-wise relative error ||x - ref|| / ||ref|| (fp64 on CPU)."""
    x = x.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-300))


def lerr(x, ref):
    """Relative error of a scalar."""
    x, ref = float(x), float(ref)
    return abs(x - ref) / max(abs(ref), 1e-300)


def planted_sent(B, D, g, frac=0.25):
    s = torch.randn(B, D, generator=g)
    k = max(1, int(B * frac))
    src = torch.randperm(B, generator=g)[:k]
    dst = torch.randperm(B, generator=g)[:k]
    for a, b in zip(src.tolist(), dst.tolist()):
        if a != b:
            s[b] = s[a] + 0.1 * torch.randn(D, generator=g)
    return s


def word_inputs(B, D, T, R, seed, lean=0.3, min_len=None):
    g = torch.Generator().manual_seed(seed)
    words = torch.randn(B, D, T, generator=g)
    regions = torch.randn(B, D, R, generator=g)
    regions = regions + lean * words[:, :, torch.randint(0, T, (R,), generator=g)]
    lo = max(1, T // 3) if min_len is None else min_len
    lens = torch.randint(lo, T + 1, (B,), generator=g)
    mask = torch.arange(T).unsqueeze(0) >= lens.unsqueeze(1)
    return words, regions, mask


_hooks_ops = None


def hooks_ops():
    """A CudaOps bound to libxmcloss_hooks.so — the -DXMC_TEST_HOOKS build of the same sources, which adds the
    debug setters (TMEM dump, NaN poisoning, generic-kernel switch).  The product library has none of them."""
    global _hooks_ops
    if _hooks_ops is None:
        from xmc_gan_b200 import _lib
        from xmc_gan_b200.ops import CudaOps
        _lib.build()
        _hooks_ops = CudaOps(lib=_lib.hooks_lib())
    return _hooks_ops
