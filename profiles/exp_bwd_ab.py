"""Same-box A/B of the backward kernel's operand hand-over in feature halves (debug flag 64 = context tile
handed over whole, flag 128 = S waits for the whole region stage): word_loss fwd+bwd at COCO-256 with the bench masks, kernel times from CUDA
events around the launches, L2 flushed between runs."""
import sys, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import _lib, train_gan as T
from xmc_gan_b200.ops import CudaOps
ops = CudaOps(lib=_lib.hooks_lib())   # -DXMC_TEST_HOOKS build: the debug flags do not exist in the product library
inp = {k: v.cuda() for k, v in bench.make_inputs(256, 1000, torch.bfloat16).items()}
labels = T.make_labels(256, inp["sent"], False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
hook = _lib.hooks_lib().xmc_internal_set_debug_dump
def run(flag, n=20):
    hook(flag)
    def step():
        v = inp["regions"].detach().requires_grad_(); w = inp["words"].detach().requires_grad_()
        loss = T.word_loss(v, w, inp["mask"], labels, False, rho1=5., rho2=5., rho3=10., precision="bf16")
        loss.backward()
        return loss, v.grad
    for _ in range(3): step()
    ops.enable_timing(True)
    for _ in range(n):
        flush.zero_()
        loss, g = step()
    k = ops.kernel_ms()
    ops.enable_timing(False)
    hook(0)
    return round(k["wordregion_bwd"][1] * 1e3, 1), round(float(loss), 6), round(float(g.float().norm()), 5)
variants = [("all_on", 0), ("whole_context_tile", 64), ("whole_region_stage", 128)]
tot = {k: 0.0 for k, _ in variants}
R = 8
for rnd in range(R):                      # rotate the order: the box drifts (clocks, temperature) within a run
    order = variants[rnd % len(variants):] + variants[:rnd % len(variants)]
    res = {k: run(f, n=10) for k, f in order}
    for k in tot: tot[k] += res[k][0]
    print({k: res[k] for k, _ in variants}, flush=True)
print("mean us:", {k: round(v / R, 1) for k, v in tot.items()})
