"""BASELINE config 5 — shape sweep: word_loss forward+backward (tcgen05 path) over batch x words x regions.
Prints one JSON line per shape: ms per fwd+bwd, kernel times, TFLOP/s on valid words (algorithmic 12*B*W*R*D), and
the parity of that very shape against the oracle's maths evaluated in fp32 by stock PyTorch on the same device
(tests/stock_losses.py, fed the bf16-rounded inputs): loss_rel and the two gradients' norm-wise errors (tol 2e-2)."""
import json, sys, torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import stock_losses
from xmc_gan_b200 import train_gan as T
from xmc_gan_b200.ops import default_ops
ops = default_ops()
D = 256
shapes = [(B, Tw, side) for B in (64, 256, 1024) for Tw in (12, 18, 32) for side in (8, 12, 16, 17)]
shapes += [(128, 18, 17), (512, 18, 17), (512, 24, 14), (1024, 18, 10)]
for B, Tw, side in shapes:
    R = side * side
    g = torch.Generator().manual_seed(B + Tw + R)
    words = torch.randn(B, D, Tw, generator=g).bfloat16().cuda()
    regions = torch.randn(B, D, side, side, generator=g).bfloat16().cuda()
    lens = torch.randint(5, Tw + 1, (B,), generator=g)
    mask = (torch.arange(Tw).unsqueeze(0) >= lens.unsqueeze(1)).cuda()
    labels = T.make_labels(B, None, False)
    def step():
        v = regions.detach().requires_grad_(); w = words.detach().requires_grad_()
        loss = T.word_loss(v, w, mask, labels, False, precision="bf16")
        loss.backward()
        return loss
    for _ in range(3): step()
    ops.enable_timing(True)
    n = 10 if B <= 256 else 4
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): l = step()
    b.record(); torch.cuda.synchronize()
    k = ops.kernel_ms(); ops.enable_timing(False)
    # parity at this shape
    v = regions.detach().requires_grad_(); w = words.detach().requires_grad_()
    lg = T.word_loss(v, w, mask, labels, False, precision="bf16"); lg.backward()
    vr = regions.float().requires_grad_(); wr = words.float().requires_grad_()
    if B <= 256:      # stock PyTorch keeps every [8, B, T, R] block alive for autograd: 40+ GB beyond B = 512
        lr = stock_losses.word_loss(vr, wr, mask, torch.eye(B, device="cuda"), False); against = "stock PyTorch fp32"
    else:             # the fp32 CUDA-core path of this library (itself held to 1e-4 against the oracle in tests/)
        lr = T.word_loss(vr, wr, mask, labels, False, precision="fp32"); against = "libxmcloss fp32 path"
    lr.backward()
    ne = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    par = {"loss_rel": abs(float(lg) - float(lr)) / abs(float(lr)), "d_regions": ne(v.grad, vr.grad), "d_words": ne(w.grad, wr.grad)}
    par["ok"] = max(par.values()) <= 2e-2
    par["against"] = against
    del vr, wr, lr
    Wv = float((~mask).sum())
    ms = a.elapsed_time(b) / n
    fl = 12.0 * B * Wv * R * D
    print(json.dumps({"B": B, "T": Tw, "R": R, "valid_words": Wv, "ms_fwd_bwd": round(ms, 4),
                      "wr_fwd_ms": round(k["wordregion_fwd"][1], 4), "wr_bwd_ms": round(k["wordregion_bwd"][1], 4),
                      "tflops_kernels": round(fl / ((k["wordregion_fwd"][1] + k["wordregion_bwd"][1]) * 1e-3) / 1e12, 1),
                      "samples_per_s": round(B / (ms * 1e-3)), "loss": round(float(l), 4),
                      "parity": {k: (round(x, 6) if isinstance(x, float) else x) for k, x in par.items()}}), flush=True)
