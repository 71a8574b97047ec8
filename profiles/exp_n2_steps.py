"""Per-step host and device times of the eager global-negative step on N ranks, with the caching
allocator's cudaMalloc count (is the start-up transient allocator growth?).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/exp_n2_steps.py [throttle]"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import train_gan as T
rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
group = dist.group.WORLD
throttle = len(sys.argv) > 1
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000 + rank, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False, group=group)
def step():
    i_ = inp["img"].detach().requires_grad_(); s_ = inp["sent"].detach().requires_grad_()
    f_ = inp["fake"].detach().requires_grad_(); w_ = inp["words"].detach().requires_grad_()
    v_ = inp["regions"].detach().requires_grad_()
    loss = (T.sent_loss(i_, s_, labels, False, group=group) + T.img_loss(inp["real"], f_, labels, False, group=group)
            + T.word_loss(v_, w_, inp["mask"], labels, False, rho1=5., rho2=5., rho3=10., precision="bf16", group=group))
    loss.backward()
N = 120
evs, host, mallocs = [], [], []
torch.cuda.synchronize(); dist.barrier()
for i in range(N):
    if throttle and i >= 2:
        evs[i - 2][1].synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(); step(); b.record()
    host.append((time.perf_counter() - t0) * 1e3)
    evs.append((a, b))
    mallocs.append(torch.cuda.memory_stats()["num_device_alloc"])
torch.cuda.synchronize()
if rank == 0:
    dev = [a.elapsed_time(b) for a, b in evs]
    for i in range(0, N, 10):
        print(f"steps {i:3d}-{i+9:3d}: host {sum(host[i:i+10])/10:6.2f} ms  device {sum(dev[i:i+10])/10:6.2f} ms  cudaMallocs so far {mallocs[i+9]}"
              f"  reserved {torch.cuda.memory_reserved() >> 20} MiB", flush=True)
dist.barrier(); dist.destroy_process_group()
