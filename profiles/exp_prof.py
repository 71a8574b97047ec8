"""One tcgen05 word-region forward + backward at COCO-256 shapes (the ncu target)."""
import sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200.ops import default_ops
ops = default_ops()
B, D, T, R = 256, 256, 18, 289
g = torch.Generator().manual_seed(0)
words = torch.randn(B, D, T, generator=g).cuda(); regions = torch.randn(B, D, R, generator=g).cuda()
qn, _ = ops.normalize_transpose(words, T, torch.bfloat16)
kn, rnorm = ops.normalize_transpose(regions, 304, torch.bfloat16)
qn = qn.view(B * T, D)
for _ in range(3):
    l, c, r, chat = ops.wordregion_forward(1, qn, kn, rnorm, R, 5.0, save_context=True)
    grel = torch.randn_like(l) * 0.1
    ops.wordregion_backward(1, qn, kn, rnorm, R, 5.0, l, c, r, grel, chat)
torch.cuda.synchronize()
print("ok")
