"""MA-GP reduction at the reference's size (B = 256, image gradients [B, 3, 256, 256] + sentence gradients
[B, 256], fp32): CUDA-event times of the fused forward / backward with the L2 flushed before each launch,
achieved GB/s against the measured HBM peak, and the reference's torch expression (train_gan.py:244-249)
on the same GPU beside it."""
import json, sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200 import train_gan as T
from xmc_gan_b200.ops import default_ops
ops = default_ops()
B = 256
gi = (torch.randn(B, 3, 256, 256, device="cuda") * 0.01).requires_grad_()
gs = (torch.randn(B, 256, device="cuda") * 0.01).requires_grad_()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    peak = 6537.6

def reference(grads):                       # the reference's expression, verbatim semantics
    grad0 = grads[0].view(grads[0].size(0), -1); grad1 = grads[1].view(grads[1].size(0), -1)
    grad = torch.cat((grad0, grad1), dim=1)
    return 2.0 * torch.mean(torch.sqrt(torch.sum(grad ** 2, dim=1)) ** 6)

def timed(fwd, n=10):
    tf = tb = 0.0
    for it in range(n + 2):
        gi.grad = None; gs.grad = None
        flush.zero_()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(); loss = fwd((gi, gs)); b.record()
        flush.zero_()
        d = torch.cuda.Event(enable_timing=True); d.record()
        loss.backward(); c.record(); torch.cuda.synchronize()
        if it >= 2:
            tf += a.elapsed_time(b); tb += d.elapsed_time(c)
    return tf / n * 1e3, tb / n * 1e3, float(loss)

nbytes = B * (3 * 256 * 256 + 256) * 4
for name, fn in (("fused (libxmcloss)", T.magp_penalty), ("torch expression", reference)):
    f, b, l = timed(fn)
    print(f"{name:20s} fwd {f:7.1f} us = {nbytes / f / 1e3:7.0f} GB/s ({nbytes / f / 1e3 / peak:.2f} of {peak:.0f})   "
          f"bwd {b:7.1f} us = {2 * nbytes / b / 1e3:7.0f} GB/s ({2 * nbytes / b / 1e3 / peak:.2f})   loss {l:.6e}")
ops.enable_timing(True)
for _ in range(5):
    gi.grad = None; gs.grad = None; flush.zero_()
    T.magp_penalty((gi, gs)).backward()
print("kernel events:", {k: round(v[1] * 1e3, 1) for k, v in ops.kernel_ms().items()}, "us (two launches in fwd)")
# The zero-fill flush leaves ~126 MB of dirty lines in L2 whose write-back shares HBM with the timed kernel;
# flushing by READING a 512 MB buffer leaves clean lines: the kernel's own traffic only.
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda").zero_()
ops.enable_timing(True)
for _ in range(8):
    gi.grad = None; gs.grad = None
    big.view(torch.int64).sum()
    loss = T.magp_penalty((gi, gs))
    big.view(torch.int64).sum()
    loss.backward()
k = ops.kernel_ms()
f, b = k["gradpen_fwd"][1] * 1e3, k["gradpen_bwd"][1] * 1e3
print(f"clean-L2 flush: fwd {f:6.1f} us = {nbytes / f / 1e3:6.0f} GB/s ({nbytes / f / 1e3 / peak:.2f})   bwd {b:6.1f} us = {2 * nbytes / b / 1e3:6.0f} GB/s ({2 * nbytes / b / 1e3 / peak:.2f})")
ops.enable_timing(False)
# slices-per-row sweep (kernel events, forward = sumsq + loss launches)
from xmc_gan_b200.ops import CudaOps
for S in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
    CudaOps._gp_slices = staticmethod(lambda B, n0, S=S: S)
    ops.enable_timing(True)
    for _ in range(6):
        gi.grad = None; gs.grad = None; flush.zero_()
        T.magp_penalty((gi, gs)).backward()
    k = ops.kernel_ms()
    print(f"slices {S:2d}: fwd {k['gradpen_fwd'][1] * 1e3:6.1f} us  bwd {k['gradpen_bwd'][1] * 1e3:6.1f} us")
