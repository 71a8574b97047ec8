"""BASELINE config 5 — shape sweep: word_loss forward+backward (tcgen05 path) over batch x words x regions.
Prints one JSON line per shape: ms per fwd+bwd, kernel times, TFLOP/s on valid words (algorithmic 12*B*W*R*D)."""
import json, sys, torch
sys.path.insert(0, '.')
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
    Wv = float((~mask).sum())
    ms = a.elapsed_time(b) / n
    fl = 12.0 * B * Wv * R * D
    print(json.dumps({"B": B, "T": Tw, "R": R, "valid_words": Wv, "ms_fwd_bwd": round(ms, 4),
                      "wr_fwd_ms": round(k["wordregion_fwd"][1], 4), "wr_bwd_ms": round(k["wordregion_bwd"][1], 4),
                      "tflops_kernels": round(fl / ((k["wordregion_fwd"][1] + k["wordregion_bwd"][1]) * 1e-3) / 1e12, 1),
                      "samples_per_s": round(B / (ms * 1e-3)), "loss": round(float(l), 4)}), flush=True)
