"""Perf experiment: clock64 timeline of the tcgen05 backward kernel (CTA 0, first chunks)."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200 import _lib
from xmc_gan_b200.ops import default_ops
ops = default_ops()
hook = _lib.lib().xmc_internal_set_debug_dump
hook.argtypes, hook.restype = [ctypes.c_int], None
B, D, T, R = 256, 256, 18, 289
g = torch.Generator().manual_seed(0)
words = torch.randn(B, D, T, generator=g).cuda(); regions = torch.randn(B, D, R, generator=g).cuda()
qn, _ = ops.normalize_transpose(words, T, torch.bfloat16)
kn, rnorm = ops.normalize_transpose(regions, 304, torch.bfloat16)
qn = qn.view(B * T, D)
l, c, r, chat = ops.wordregion_forward(1, qn, kn, rnorm, R, 5.0, save_context=True)
grel = torch.randn_like(l) * 0.1
for _ in range(2):
    ops.wordregion_backward(1, qn, kn, rnorm, R, 5.0, l, c, r, grel, chat)
hook(4)
ops.wordregion_backward(1, qn, kn, rnorm, R, 5.0, l, c, r, grel, chat)
torch.cuda.synchronize()
hook(0)
tr = ops.last_workspace[64:64 + 4 * 64 * 4 * 8].view(torch.int64).view(4, 64, 4).cpu()
t0 = int(tr[1, 0, 0])
print("chunk | MMA: xy-ready, B issued, kv-ready, A issued | EW: sw-ready, E done, dk-ready, drain done   (cycles since first sw-ready)")
for g_ in range(5, 21):
    m = [int(v) - t0 for v in tr[0, g_]]
    e = [int(v) - t0 for v in tr[1, g_]]
    x = [int(v) for v in tr[2, g_]]
    y = [int(v) for v in tr[3, g_]]
    print(f"      drain quarter 2: start@{y[0]-int(tr[1,g_,2])} after dk-ready; wait_read+bar={y[1]-y[0]} ld+sts+fence={y[2]-y[1]} bar={y[3]-y[2]}")
    print(f"      E parts: ld={x[0]-int(tr[1,g_,0])} math={x[1]-x[0]} wait+bar={x[2]-x[1]} stores+colsum={x[3]-x[2]}")
    print(f"{g_:3d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} | {e[0]:7d} {e[1]:7d} {e[2]:7d} {e[3]:7d} | E={e[1]-e[0]:5d} waitdk={e[2]-e[1]:5d} drain={e[3]-e[2]:5d} period={int(tr[1,g_,0])-int(tr[1,g_-1,0]):6d}")
