"""fp32-tolerance word-region kernels (split-bf16 tcgen05, path 2) at COCO-256 with the bench masks: CUDA-event time of
the forward and backward launches, next to the fp32 CUDA-core kernels (path 0) on the same operands."""
import json, sys, torch
sys.path.insert(0, '.')
import bench
from xmc_gan_b200 import _lib
from xmc_gan_b200.ops import default_ops
ops = default_ops()
B, D, T, R = 256, 256, 18, 289
inp = bench.make_inputs(B, 1000)
mask = inp["mask"].to(torch.uint8).cuda()
row_of, cap_ptr = ops.word_rows_compact(mask)
nq = cap_ptr[B:]
qn, _ = ops.normalize_transpose(inp["words"].cuda(), T, torch.float32, row_of=row_of)
kn, rnorm = ops.normalize_transpose(inp["regions"].flatten(2).cuda(), 304, torch.float32)
qn = qn.view(B * T, D)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {}
for name, path in (("split_tcgen05", _lib.PATH_FP32_TCGEN05), ("cuda_core", _lib.PATH_FP32_SIMT)):
    l, c, r, chat = ops.wordregion_forward(path, qn, kn, rnorm, R, 5.0, save_context=True, nq_dev=nq)
    grel = torch.randn_like(l) * 0.01
    out[name + "_fwd_ms"] = round(timed(lambda: ops.wordregion_forward(path, qn, kn, rnorm, R, 5.0, save_context=True, nq_dev=nq)), 3)
    out[name + "_bwd_ms"] = round(timed(lambda: ops.wordregion_backward(path, qn, kn, rnorm, R, 5.0, l, c, r, grel, chat, nq_dev=nq)), 3)
valid = float((mask == 0).sum())
out["valid_words"] = valid
out["algorithmic_TFLOPs_bwd_split"] = round(8.0 * B * valid * R * D / (out["split_tcgen05_bwd_ms"] * 1e-3) / 1e12, 1)
out["algorithmic_TFLOPs_fwd_split"] = round(4.0 * B * valid * R * D / (out["split_tcgen05_fwd_ms"] * 1e-3) / 1e12, 1)
print(json.dumps(out))
# phase breakdown of the backward (hooks build): cycles of CTA (0, 0) per phase, summed over its chunks
from xmc_gan_b200.ops import CudaOps
hops = CudaOps(lib=_lib.hooks_lib())
l, c, r, chat = hops.wordregion_forward(_lib.PATH_FP32_TCGEN05, qn, kn, rnorm, R, 5.0, save_context=True, nq_dev=nq)
hops.wordregion_backward(_lib.PATH_FP32_TCGEN05, qn, kn, rnorm, R, 5.0, l, c, r, torch.randn_like(l) * 0.01, chat, nq_dev=nq)
torch.cuda.synchronize()
ph = hops.last_workspace[16:16 + 8 * 10].view(torch.int64).tolist()
names = ["stage regions", "stage Q halves (S)", "S MMAs", "stage C halves (W)", "W MMAs", "elementwise X, Y", "dQ + dK[1] C-part MMAs",
         "stage halves (dK)", "dK MMAs", "dK drains"]
tot = sum(ph)
print(json.dumps({"backward_phase_cycles_cta0": {n: v for n, v in zip(names, ph)}, "total": tot,
                  "chunks": 5 * len(range(0, B, max(1, (B + 3) // 4)))}))
