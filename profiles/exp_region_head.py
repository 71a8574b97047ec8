"""Region head of the word loss at BASELINE config 4's shape: the fused tcgen05 kernels of
region_head.cu against the unfused chain (PyTorch 1x1 convolution + layout/normalise kernel, and the convolution's
cuDNN/cuBLAS backward); CUDA events, L2 flushed between launches."""
import json, sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
import os
from xmc_gan_b200.ops import CudaOps, default_ops
from xmc_gan_b200 import _lib
# XMC_HEAD_PAIR=1: the 2-SM form of the kernels, compiled into the hooks build only
ops = CudaOps(lib=_lib.hooks_lib()) if os.environ.get("XMC_HEAD_PAIR") else default_ops()
# L2 is flushed by READING a 512 MB buffer: a zero-fill leaves ~126 MB of dirty lines whose write-back is then charged to
# the timed kernel (measured: +15-20 us on these HBM-bound kernels, enough to hide every difference between variants)
flush = torch.zeros(512 << 18, dtype=torch.float32, device="cuda")
PEAK = 6537.6   # GB/s, MEASURED_PEAKS.json


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2] * 1e3


B, Cin, H, W, D = 256, 512, 16, 16, 256
R = H * W
for fdt in (torch.float32, torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(1)
    feat = torch.randn(B, Cin, H, W, generator=g, device="cuda").to(fdt)
    w = (torch.randn(D, Cin, generator=g, device="cuda") / Cin ** 0.5).to(fdt)
    bias = torch.randn(D, generator=g, device="cuda") * 0.1
    dy = (torch.randn(B, R, D, generator=g, device="cuda") * 0.01).to(fdt)     # the loss hands dy over in the map's dtype
    s = feat.element_size()
    f3 = feat.flatten(2)

    def unfused_fwd():
        y = F.conv2d(feat, w.view(D, Cin, 1, 1), bias.to(fdt))                  # [B, D, H, W]
        return ops.normalize_transpose(y.flatten(2), R, torch.bfloat16)

    def unfused_bwd():
        dyc = dy.transpose(1, 2).reshape(B, D, H, W).to(fdt)
        gi = torch.ops.aten.convolution_backward(dyc, feat, w.view(D, Cin, 1, 1), [D], [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
                                                 [True, True, True])
        return gi

    res = {"feat_dtype": str(fdt).split(".")[-1]}
    res["fwd_fused_us"] = round(timed(lambda: ops.region_head_forward(f3, w, bias, R)), 1)
    res["fwd_unfused_us"] = round(timed(unfused_fwd), 1)
    res["bwd_input_us"] = round(timed(lambda: ops.region_head_backward(f3, w, dy, True, False, False)), 1)
    res["bwd_weight_us"] = round(timed(lambda: ops.region_head_backward(f3, w, dy, False, True, True)), 1)
    res["bwd_unfused_us"] = round(timed(unfused_bwd), 1)
    by_f = B * R * (Cin * s + D * 2 + 4)
    by_i = B * R * (D * s + Cin * s)
    by_w = B * R * (D * s + Cin * s)
    res.update(algorithmic_bytes_fwd=by_f, fwd_GBs=round(by_f / res["fwd_fused_us"] / 1e3), fwd_frac=round(by_f / res["fwd_fused_us"] / 1e3 / PEAK, 3),
               bwd_input_GBs=round(by_i / res["bwd_input_us"] / 1e3), bwd_input_frac=round(by_i / res["bwd_input_us"] / 1e3 / PEAK, 3),
               bwd_weight_GBs=round(by_w / res["bwd_weight_us"] / 1e3), bwd_weight_frac=round(by_w / res["bwd_weight_us"] / 1e3 / PEAK, 3),
               flops_each=2 * B * R * Cin * D)
    print(json.dumps({"B": B, "Cin": Cin, "R": R, "D": D, **res}), flush=True)
