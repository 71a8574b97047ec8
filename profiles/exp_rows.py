"""Region prologue / epilogue of the word loss at COCO shapes: row-layout kernels on a channels-last map
(xmc_normalize_rows) vs the transposing kernels on the reference layout; CUDA events, L2 flushed."""
import json, sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200.ops import default_ops
ops = default_ops()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2] * 1e3


for B, D, side in ((256, 256, 16), (256, 256, 17)):
    R, Rpad = side * side, (side * side + 15) // 16 * 16
    g = torch.Generator().manual_seed(side)
    x = torch.randn(B, D, R, generator=g).bfloat16().cuda()
    rows = x.transpose(1, 2).contiguous()
    dkn = torch.randn(B, Rpad, D, generator=g).cuda()
    drn = torch.randn(B, Rpad, generator=g).cuda()
    kn, rn = ops.normalize_transpose(x, Rpad, torch.bfloat16)
    res = {
        "fwd_transpose_us": round(timed(lambda: ops.normalize_transpose(x, Rpad, torch.bfloat16)), 1),
        "fwd_rows_us": round(timed(lambda: ops.normalize_rows(rows, Rpad, torch.bfloat16)), 1),
        "bwd_transpose_us": round(timed(lambda: ops.normalize_transpose_backward(kn, rn, dkn, drn, R, torch.bfloat16)), 1),
        "bwd_rows_us": round(timed(lambda: ops.normalize_rows_backward(kn, rn, dkn, drn, R, torch.bfloat16)), 1),
    }
    by_f = B * R * D * 2 * 2
    by_b = B * R * D * (2 + 4 + 2)
    res["fwd_rows_GBs"] = round(by_f / res["fwd_rows_us"] / 1e3, 0)
    res["bwd_rows_GBs"] = round(by_b / res["bwd_rows_us"] / 1e3, 0)
    print(json.dumps({"B": B, "D": D, "R": R, **res, "algorithmic_bytes_fwd": by_f, "algorithmic_bytes_bwd": by_b}), flush=True)
