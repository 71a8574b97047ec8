"""GPU timeline of the global-negative step on N ranks (torchrun): kernel/collective intervals of a few
eager steps from torch.profiler (CUPTI), with the idle gaps between them on rank 0.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/exp_n2_timeline.py"""
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import train_gan as T
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
group = dist.group.WORLD
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000 + rank, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False, group=group)
def step():
    i_ = inp["img"].detach().requires_grad_(); s_ = inp["sent"].detach().requires_grad_()
    f_ = inp["fake"].detach().requires_grad_(); w_ = inp["words"].detach().requires_grad_()
    v_ = inp["regions"].detach().requires_grad_()
    loss = (T.sent_loss(i_, s_, labels, False, group=group) + T.img_loss(inp["real"], f_, labels, False, group=group)
            + T.word_loss(v_, w_, inp["mask"], labels, False, rho1=5., rho2=5., rho3=10., precision="bf16", group=group))
    loss.backward()
for _ in range(10): step()
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4): step()
    torch.cuda.synchronize()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    prof.export_chrome_trace("gpurun_out/n2_trace.json")
    ev = [e for e in json.load(open("gpurun_out/n2_trace.json"))["traceEvents"]
          if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]; end = max(e["ts"] + e["dur"] for e in ev)
    # one step = between consecutive wr_bwd kernels
    b = [i for i, e in enumerate(ev) if "wr_bwd" in e["name"]]
    lo, hi = b[-2] + 1, b[-1] + 1
    step_ev = ev[lo:hi]
    s0 = ev[b[-2]]["ts"] + ev[b[-2]]["dur"]
    print(f"span of one step (bwd end -> bwd end): {ev[b[-1]]['ts'] + ev[b[-1]]['dur'] - s0:.0f} us, {len(step_ev)} GPU activities")
    cur = s0
    for e in step_ev:
        gap = e["ts"] - cur
        print(f"{e['ts'] - s0:8.0f} us  +{e['dur']:7.1f}  gap {gap:7.1f}  stream {e['args'].get('stream')}  {e['name'][:70]}")
        cur = max(cur, e["ts"] + e["dur"])
dist.barrier(); dist.destroy_process_group()
