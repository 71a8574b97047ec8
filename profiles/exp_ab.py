"""Same-box A/B of the small-kernel changes around the word-region kernels: word_loss fwd+bwd at COCO-256,
captured as a CUDA graph per variant, replays timed with CUDA events (L2 flushed between replays),
variants interleaved over several rounds."""
import sys, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import _lib, train_gan as T
from xmc_gan_b200.ops import CudaOps
ops = CudaOps(lib=_lib.hooks_lib())   # -DXMC_TEST_HOOKS build: the A/B switches do not exist in the product library
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
v = inp["regions"].detach().requires_grad_(); w = inp["words"].detach().requires_grad_()

def step():
    v.grad = None; w.grad = None
    loss = T.word_loss(v, w, inp["mask"], labels, False, rho1=5., rho2=5., rho3=10., precision="bf16")
    loss.backward()
    return loss

def capture():
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    return g, loss

fused = CudaOps.word_scores_infonce_backward
def variant(name):
    ops.use_side_stream = name != "no_side_stream"
    _lib.hooks_lib().xmc_internal_set_prep_generic(int(name == "generic_prep"))
    if name == "two_call_tail":
        del CudaOps.word_scores_infonce_backward
    try:
        return capture()
    finally:
        CudaOps.word_scores_infonce_backward = fused
        ops.use_side_stream = True
        _lib.hooks_lib().xmc_internal_set_prep_generic(0)

names = ["all_on", "no_side_stream", "two_call_tail", "generic_prep"]
graphs = {n: variant(n) for n in names}
def timed(g, n=30):
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3
for n in names:
    for _ in range(5): graphs[n][0].replay()
for rnd in range(3):
    print({n: round(timed(graphs[n][0]), 1) for n in names}, "us per word_loss fwd+bwd", flush=True)
print({n: round(float(graphs[n][1]), 6) for n in names})
