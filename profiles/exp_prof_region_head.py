"""ncu target: the region head of the word loss at BASELINE config 4's shape (B = 256, [512, 16, 16] map, D = 256), bf16 and
fp32 (tf32) operands: forward, dfeat, dW + dbias."""
import sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200.ops import default_ops
ops = default_ops()
g = torch.Generator(device="cuda").manual_seed(0)
B, Cin, R, D = 256, 512, 256, 256
for dt in (torch.bfloat16, torch.float32):
    feat = torch.randn(B, Cin, R, generator=g, device="cuda").to(dt)
    w = (torch.randn(D, Cin, generator=g, device="cuda") / Cin ** 0.5).to(dt)
    bias = torch.randn(D, generator=g, device="cuda") * 0.1
    dy = (torch.randn(B, R, D, generator=g, device="cuda") * 0.01).to(dt)
    for _ in range(2):
        ops.region_head_forward(feat, w, bias, R)
        ops.region_head_backward(feat, w, dy, True, True, True)
torch.cuda.synchronize()
print("ok")
