"""Perf experiment: clock64 timeline of the tcgen05 backward kernel (CTA 0, first chunks)."""
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
print("chunk | MMA: SW(g+1) issue start, XY(g) ready, dQ/dK(g) issued, end | arith(g): sw ready, consumed, done | loop(g): XY stored, arith(g+1) done, dk ready, drain done")
for g_ in range(1, 60):
    m = [int(v) - t0 for v in tr[0, g_]]
    a = [int(v) - t0 for v in tr[1, g_]]
    e = [int(v) - t0 for v in tr[2, g_]]
    print(f"{g_:3d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} | {a[0]:7d} {a[1]:7d} {a[2]:7d} (ld {a[1]-a[0]}, math {a[2]-a[1]}) | {e[0]:7d} {e[1]:7d} {e[2]:7d} {e[3]:7d} (wait dk {e[2]-e[1]}, drain {e[3]-e[2]}) period {e[0]-int(tr[2,g_-1,0])+t0}")
