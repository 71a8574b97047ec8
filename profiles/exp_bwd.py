import ctypes, sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200 import _lib
from xmc_gan_b200.ops import CudaOps
ops = CudaOps(lib=_lib.hooks_lib())   # -DXMC_TEST_HOOKS build: the debug flags do not exist in the product library
hook = _lib.hooks_lib().xmc_internal_set_debug_dump
B, D, T, R = 256, 256, 18, 289
g = torch.Generator().manual_seed(0)
words = torch.randn(B, D, T, generator=g).cuda(); regions = torch.randn(B, D, R, generator=g).cuda()
qn, _ = ops.normalize_transpose(words, T, torch.bfloat16)
kn, rnorm = ops.normalize_transpose(regions, 304, torch.bfloat16)
qn = qn.view(B * T, D)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for rn in (rnorm, None):
    l, c, r, chat = ops.wordregion_forward(1, qn, kn, rn, R, 5.0, save_context=True)
    grel = torch.randn_like(l) * 0.1
    for flags in (0, 2):
        hook(flags)
        ms = timeit(lambda: ops.wordregion_backward(1, qn, kn, rn, R, 5.0, l, c, r, grel, chat))
        print(f"raw_values={rn is not None} flags={flags}: bwd {ms:.3f} ms (includes 3 zero-fills)")
    hook(0)
    ms = timeit(lambda: ops.wordregion_forward(1, qn, kn, rn, R, 5.0, save_context=True))
    ms2 = timeit(lambda: ops.wordregion_forward(1, qn, kn, rn, R, 5.0, save_context=False))
    print(f"raw_values={rn is not None}: fwd {ms:.3f} ms with chat, {ms2:.3f} ms without")
