"""ncu target: word_loss fwd+bwd at COCO-256 with the bench masks, fp32 inputs -> the split-bf16 tcgen05 kernels (fp32 tolerance)."""
import sys, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import train_gan as T
inp = {k: (v.cuda().float() if v.dtype.is_floating_point else v.cuda()) for k, v in bench.make_inputs(256, 1000, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False)
for _ in range(3):
    v = inp["regions"].detach().requires_grad_(); w = inp["words"].detach().requires_grad_()
    loss = T.word_loss(v, w, inp["mask"], labels, False, rho1=5., rho2=5., rho3=10., precision="fp32")
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss.detach()))
