"""Where the host time of one eagerly launched step goes (cProfile over 100 steps at COCO-256, one GPU)."""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import train_gan as T
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False)
def step():
    i_ = inp["img"].detach().requires_grad_(); s_ = inp["sent"].detach().requires_grad_()
    f_ = inp["fake"].detach().requires_grad_(); w_ = inp["words"].detach().requires_grad_()
    v_ = inp["regions"].detach().requires_grad_()
    loss = (T.sent_loss(i_, s_, labels, False) + T.img_loss(inp["real"], f_, labels, False)
            + T.word_loss(v_, w_, inp["mask"], labels, False, rho1=5., rho2=5., rho3=10., precision="bf16"))
    loss.backward()
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100): step()
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"host enqueue {1e3 * (t1 - t0) / 100:.3f} ms per step")
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
