"""Perf experiment: clock64 timeline of the tcgen05 forward kernel (CTA 0, first 64 chunks / images)."""
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
for _ in range(2):
    ops.wordregion_forward(1, qn, kn, rnorm, R, 5.0, save_context=True)
hook(4 | (int(sys.argv[1]) if len(sys.argv) > 1 else 0))
ops.wordregion_forward(1, qn, kn, rnorm, R, 5.0, save_context=True)
torch.cuda.synchronize()
hook(0)
tr = ops.last_workspace[64:64 + 4 * 64 * 4 * 8].view(torch.int64).view(4, 64, 4).cpu()
t0 = int(tr[1, 0, 0])
print("chunk | MMA: P(g)+kv(g+2) ready, G2(g) issued, G1(g+2) issued | SM: S ready, ld done, math+st done, arrived")
for g_ in range(0, 32):
    m = [int(v) - t0 for v in tr[0, g_]]
    e = [int(v) - t0 for v in tr[1, g_]]
    per = e[0] - (int(tr[1, g_ - 1, 0]) - t0) if g_ else 0
    print(f"{g_:3d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} | {e[0]:7d} {e[1]:7d} {e[2]:7d} {e[3]:7d} | ld={e[1]-e[0]:4d} math={e[2]-e[1]:4d} arr={e[3]-e[2]:4d} period={per:5d} | exp={int(tr[3,g_,0])-int(tr[1,g_,1])} rest={int(tr[3,g_,1])-int(tr[3,g_,0])} st={int(tr[1,g_,2])-int(tr[3,g_,1])}")
print("image | c_full ready, C in registers (c_empty), stores issued")
for ii in range(0, 6):
    e = [int(v) - t0 for v in tr[2, ii]]
    print(f"{ii:3d} | {e[0]:7d} {e[1]:7d} {e[3]:7d} | read={e[1]-e[0]} rest={e[3]-e[1]}")
clk, ns, nimg = (int(v) for v in tr[2, 63, :3])
print(f"CTA 0 main loop: {clk} SM cycles in {ns} ns over {nimg} images = {clk / max(ns, 1) * 1e3:.0f} MHz effective SM clock")
