"""Perf experiment: per-CTA wall time of the tcgen05 backward kernel (debug flag 8): start/end ns, segments, images."""
import ctypes, sys, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import _lib
from xmc_gan_b200.ops import CudaOps
ops = CudaOps(lib=_lib.hooks_lib())   # -DXMC_TEST_HOOKS build: the debug flags do not exist in the product library
hook = _lib.hooks_lib().xmc_internal_set_debug_dump
B, D, T, R = 256, 256, 18, 289
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000, torch.bfloat16).items()}
for compact in (True, False):
    mask = inp["mask"].to(torch.uint8)
    if compact:
        row_of, cap_ptr = ops.word_rows_compact(mask); nq = cap_ptr[B:]
        qn, _ = ops.normalize_transpose(inp["words"], T, torch.bfloat16, row_of=row_of)
    else:
        nq = None
        qn, _ = ops.normalize_transpose(inp["words"], T, torch.bfloat16)
    kn, rnorm = ops.normalize_transpose(inp["regions"].flatten(2), 304, torch.bfloat16)
    qn = qn.view(B * T, D)
    l, c, r, chat = ops.wordregion_forward(1, qn, kn, rnorm, R, 5.0, save_context=True, nq_dev=nq)
    grel = torch.randn_like(l) * 0.1
    for _ in range(2):
        ops.wordregion_backward(1, qn, kn, rnorm, R, 5.0, l, c, r, grel, chat, nq_dev=nq)
    hook(8)
    ops.wordregion_backward(1, qn, kn, rnorm, R, 5.0, l, c, r, grel, chat, nq_dev=nq)
    torch.cuda.synchronize()
    hook(0)
    base = 64 + 4 * 64 * 4 * 8
    log = ops.last_workspace[base:base + 148 * 4 * 8].view(torch.int64).view(148, 4).cpu()
    t0 = int(log[:, 0].min())
    dur = (log[:, 1] - log[:, 0]).float() / 1000.0
    print(f"compact={compact}: kernel span {(int(log[:,1].max()) - t0)/1000:.1f} us; per-CTA duration us: min {dur.min():.1f} mean {dur.mean():.1f} max {dur.max():.1f}")
    per_img = dur / log[:, 3].float()
    print("   us per image: min %.2f mean %.2f max %.2f; images per CTA min %d max %d; segments: %s" % (
        per_img.min(), per_img.mean(), per_img.max(), int(log[:, 3].min()), int(log[:, 3].max()),
        {int(k): int((log[:, 2] == k).sum()) for k in log[:, 2].unique()}))
    for nseg in log[:, 2].unique():
        m = log[:, 2] == nseg
        print(f"   CTAs with {int(nseg)} segments: mean duration {dur[m].mean():.1f} us, mean images {log[m, 3].float().mean():.1f}, us/image {per_img[m].mean():.2f}")
    order = torch.argsort(dur, descending=True)[:5]
    print("   slowest CTAs:", [(int(i), round(float(dur[i]), 1), int(log[i, 2]), int(log[i, 3])) for i in order])
