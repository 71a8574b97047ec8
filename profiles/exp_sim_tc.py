"""Similarity losses at the rank-shaped problem of 8-GPU global negatives (256 local rows x 2048 gathered columns):
tcgen05 form (simloss_tc.cu) vs the CUDA-core form on the same inputs; CUDA-event time per launch group, L2 flushed.
Also the ncu target for the sim_tc_* kernels (`-k regex:sim_tc_`)."""
import json, sys, torch
sys.path.insert(0, '.')
from xmc_gan_b200.ops import default_ops
ops = default_ops()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2] * 1e3      # median, us


for dt in (torch.bfloat16, torch.float32):
    for Bq, Bk, D in ((256, 2048, 256), (256, 2048, 512), (256, 512, 256), (256, 1024, 512)):
        g = torch.Generator().manual_seed(D)
        a = torch.randn(Bq, D, generator=g).to(dt).cuda()
        b = torch.randn(Bk, D, generator=g).to(dt).cuda()
        go = torch.ones((), device="cuda")
        out = {}
        for name, tc in (("tcgen05", True), ("cuda_core", False)):
            ops.use_sim_tc = tc
            sc, ia, ib, rs, cs = ops.simloss_forward(a, b, None, 0, 1.0)
            bwd = lambda: ops.simloss_backward(a, b, sc, ia, ib, None, 0, 1.0, rs, cs, None, None, 1.0, Bq, Bk, go, True, True)
            out[name + "_bwd_us"] = round(timed(bwd), 1)
        ops.use_sim_tc = True
        out["fwd_us (tcgen05 tiles + statistics pass)"] = round(timed(lambda: ops.simloss_forward(a, b, None, 0, 1.0)), 1)
        flops = 2.0 * Bq * Bk * D
        print(json.dumps({"dtype": str(dt).split(".")[1], "Bq": Bq, "Bk": Bk, "D": D, **out,
                          "algorithmic_MFLOP_per_product": round(flops / 1e6, 1),
                          "algorithmic_bytes_bwd": (Bq + Bk) * D * (a.element_size() * 2) + Bq * Bk * 4}), flush=True)
