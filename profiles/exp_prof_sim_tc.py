"""ncu target: sent_loss-shaped similarity loss at the 8-GPU rank problem (256 x 2048, D = 256, bf16), tcgen05 form."""
import sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200.ops import default_ops
ops = default_ops()
g = torch.Generator().manual_seed(0)
a = torch.randn(256, 256, generator=g).bfloat16().cuda()
b = torch.randn(2048, 256, generator=g).bfloat16().cuda()
go = torch.ones((), device="cuda")
for _ in range(3):
    sc, ia, ib, rs, cs = ops.simloss_forward(a, b, None, 0, 1.0)
    ops.simloss_backward(a, b, sc, ia, ib, None, 0, 1.0, rs, cs, None, None, 1.0, 256, 2048, go, True, True)
torch.cuda.synchronize()
print("ok")
