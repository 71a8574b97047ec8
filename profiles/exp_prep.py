"""A/B timing of the layout kernels (xmc_normalize_transpose and its backward) at COCO-256 shapes:
generic kernels vs the bf16 fast kernels, L2 flushed before every timed launch."""
import sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200 import _lib
from xmc_gan_b200.ops import CudaOps
ops = CudaOps(lib=_lib.hooks_lib())   # -DXMC_TEST_HOOKS build: the debug flags do not exist in the product library
B, D, R, Rpad, T = 256, 256, 289, 304, 18
reg = torch.randn(B, D, R, device="cuda").bfloat16()
words = torch.randn(B, D, T, device="cuda").bfloat16()
mask = (torch.arange(T, device="cuda")[None] >= torch.randint(5, T + 1, (B, 1), device="cuda")).to(torch.uint8)
row_of, _ = ops.word_rows_compact(mask)
dkn = torch.randn(B, Rpad, D, device="cuda")
drn = torch.randn(B, Rpad, device="cuda")
dqn = torch.randn(B, T, D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, n=20):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3

for generic in (1, 0, 1, 0):
    _lib.hooks_lib().xmc_internal_set_prep_generic(generic)
    kn, rnorm = ops.normalize_transpose(reg, Rpad, torch.bfloat16)
    qn, qnorm = ops.normalize_transpose(words, T, torch.bfloat16, row_of=row_of)
    t = dict(
        fwd_regions=timed(lambda: ops.normalize_transpose(reg, Rpad, torch.bfloat16)),
        fwd_words=timed(lambda: ops.normalize_transpose(words, T, torch.bfloat16, row_of=row_of)),
        bwd_regions=timed(lambda: ops.normalize_transpose_backward(kn, rnorm, dkn, drn, R, torch.bfloat16)),
        bwd_words=timed(lambda: ops.normalize_transpose_backward(qn, qnorm, dqn, None, T, torch.float32, row_of=row_of)),
    )
    print("generic" if generic else "fast   ", {k: round(v, 1) for k, v in t.items()}, "us", flush=True)
_lib.hooks_lib().xmc_internal_set_prep_generic(0)
