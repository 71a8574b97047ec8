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
