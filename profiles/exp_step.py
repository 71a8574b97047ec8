"""Perf experiment: word_loss fwd+bwd through the public API at COCO-256 shapes (masks as in bench.py)."""
import sys, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import train_gan as T
from xmc_gan_b200.ops import default_ops
ops = default_ops()
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False)
def step():
    v = inp["regions"].detach().requires_grad_(); w = inp["words"].detach().requires_grad_()
    loss = T.word_loss(v, w, inp["mask"], labels, False, rho1=5., rho2=5., rho3=10., precision="bf16")
    loss.backward()
    return loss
for compact in (True, False):
    ops.supports_compaction = compact
    for _ in range(3): step()
    ops.enable_timing(True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): l = step()
    b.record(); torch.cuda.synchronize()
    print(f"compact={compact}: word_loss fwd+bwd {a.elapsed_time(b)/10:.3f} ms/step, loss {float(l):.5f}, kernels {ops.kernel_ms()}")
    ops.enable_timing(False)
print("valid word rows:", int((~inp["mask"]).sum()), "of", inp["mask"].numel())
