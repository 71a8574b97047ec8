"""ncu target: sent_loss / img_loss fwd+bwd at COCO-256 shapes (bf16)."""
import sys, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import train_gan as T
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False)
for _ in range(3):
    i_ = inp["img"].detach().requires_grad_(); s_ = inp["sent"].detach().requires_grad_(); f_ = inp["fake"].detach().requires_grad_()
    loss = T.sent_loss(i_, s_, labels, False) + T.img_loss(inp["real"], f_, labels, False)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
