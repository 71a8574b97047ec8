"""Generate the committed golden vectors under tests/golden/.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

* ``ref_*.npz``  — inputs and outputs of the REFERENCE's own functions
  (``xmc_gan/train_gan.py:72-139``, AST-loaded unmodified, CPU fp32 and fp64).
  These pin the oracle restatement and, on the GPU box, the CUDA path.
* ``ref_attn_*.npz`` — the reference's attention block (``xmc_gan/model/concept_gan.py:532-555``, AST-loaded) on
  (image, caption) pairs: pins the attention stage of the word–region restatement (``oracle.attend``).
* ``word_*.npz`` — outputs of this repo's word–region restatement in float64
  (PARITY UNPINNED: the reference has no word loss); they only guard against
  regressions of the oracle itself.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import load_reference as LR  # noqa: E402
from oracle import word_region as WR  # noqa: E402


def planted(B, D, g, frac=0.25):
    """Sentence embeddings with near-duplicate pairs so cos > 0.6 fires."""
    s = torch.randn(B, D, generator=g)
    k = max(1, int(B * frac))
    src = torch.randperm(B, generator=g)[:k]
    dst = torch.randperm(B, generator=g)[:k]
    for a, b in zip(src.tolist(), dst.tolist()):
        if a != b:
            s[b] = s[a] + 0.1 * torch.randn(D, generator=g)
    return s


def run_ref(fn, a, b, labels, b_global, dtype, need=(True, True)):
    a = a.to(dtype).clone().requires_grad_(need[0])
    b = b.to(dtype).clone().requires_grad_(need[1])
    loss = fn(a, b, labels, b_global)
    loss.backward()
    z = lambda t, r: (t.grad if r else torch.zeros_like(t)).detach().numpy()
    return loss.detach().numpy(), z(a, need[0]), z(b, need[1])


def sim_case(name, kind, B, D, b_global, smooth, seed, need=(True, True)):
    g = torch.Generator().manual_seed(seed)
    ref = LR.load_reference_losses(smooth)
    a = torch.randn(B, D, generator=g)
    b = torch.randn(B, D, generator=g) + 0.5 * a        # correlated so the diagonal matters
    sent = planted(B, 48, g)
    with LR.cuda_is_identity():
        labels = ref.make_labels(B, sent, b_global)
    fn = ref.sent_loss if kind == "sent" else ref.img_loss
    l32, da32, db32 = run_ref(fn, a, b, labels, b_global, torch.float32, need)
    l64, da64, db64 = run_ref(fn, a, b, labels, b_global, torch.float64, need)
    sc = ref.cosine_scores(a, b).numpy()
    np.savez_compressed(
        os.path.join(HERE, f"ref_{name}.npz"),
        kind=kind, b_global=b_global, smooth_global=smooth, need=np.array(need),
        a=a.numpy(), b=b.numpy(), sent=sent.numpy(), labels=labels.numpy(), scores=sc,
        loss32=l32, da32=da32, db32=db32, loss64=l64, da64=da64, db64=db64)
    print(f"ref_{name}: loss={float(l64):.12f}")


def word_case(name, B, D, T, R, seed, rho=(5.0, 5.0, 10.0), normalize_values=False,
              b_global=False, smooth=0.5, lean=0.7):
    g = torch.Generator().manual_seed(seed)
    words = torch.randn(B, D, T, generator=g, dtype=torch.float64)
    regions = torch.randn(B, D, R, generator=g, dtype=torch.float64)
    # make the matching pair informative: image i's regions lean towards caption i's words
    regions = regions + lean * words[:, :, torch.randint(0, T, (R,), generator=g)]
    lens = torch.randint(max(1, T // 3), T + 1, (B,), generator=g)
    mask = torch.arange(T).unsqueeze(0) >= lens.unsqueeze(1)
    if B > 4:
        mask[3] = True                                  # one fully padded caption
    sent = planted(B, 48, g)
    labels = __import__("oracle").make_labels(B, sent, b_global, smooth_global=smooth)
    w = words.clone().requires_grad_(True)
    r = regions.clone().requires_grad_(True)
    S = WR.word_scores(r, w, mask, rho[0], rho[1], normalize_values)
    loss = WR.word_loss(r, w, mask, labels, b_global, smooth, *rho, normalize_values)
    loss.backward()
    np.savez_compressed(
        os.path.join(HERE, f"word_{name}.npz"),
        words=words.numpy().astype(np.float32), regions=regions.numpy().astype(np.float32),
        mask=mask.numpy(), labels=labels.numpy(), rho=np.array(rho),
        normalize_values=normalize_values, b_global=b_global, smooth_global=smooth,
        scores=S.detach().numpy(), loss=loss.detach().numpy(),
        dwords=w.grad.numpy().astype(np.float32), dregions=r.grad.numpy().astype(np.float32))
    print(f"word_{name}: loss={float(loss.detach()):.12f}")


def magp_case(name, B, shape_img, D, seed, scale):
    """The reference's own six statements (train_gan.py:244-249) on synthetic gradients."""
    g = torch.Generator().manual_seed(seed)
    ref = LR.load_reference_magp()
    gi = torch.randn(B, *shape_img, generator=g) * scale
    gs = torch.randn(B, D, generator=g) * scale
    out = {}
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        a = gi.to(dt).clone().requires_grad_(); b = gs.to(dt).clone().requires_grad_()
        loss = ref((a, b))
        loss.backward()
        out["loss" + tag] = loss.detach().numpy(); out["d0_" + tag] = a.grad.numpy(); out["d1_" + tag] = b.grad.numpy()
    np.savez_compressed(os.path.join(HERE, f"ref_magp_{name}.npz"), g0=gi.numpy(), g1=gs.numpy(), **out)
    print(f"ref_magp_{name}: loss={float(out['loss64']):.12f}")


def attn_case(name, Bi, Bc, D, T, R, seed):
    """The reference's attention block (concept_gan.py:532-555, AST-loaded) on every (image, caption) pair: queries = the
    caption's words [D, T], keys = the image's regions [D, R], no key masked -> contexts [Bi, Bc, T, D] (sum of unit keys)."""
    g = torch.Generator().manual_seed(seed)
    ref = LR.load_reference_attention()
    words = torch.randn(Bc, D, T, generator=g)
    regions = torch.randn(Bi, D, R, generator=g) + 0.5 * words[torch.arange(Bi) % Bc][:, :, torch.randint(0, T, (R,), generator=g)]
    out = {}
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        q = words.to(dt).unsqueeze(0).expand(Bi, Bc, D, T).reshape(Bi * Bc, D, T)
        k = regions.to(dt).unsqueeze(1).expand(Bi, Bc, D, R).reshape(Bi * Bc, D, R)
        mask = torch.zeros(Bi * Bc, R, dtype=torch.bool)
        ctx = ref(q.clone(), k.clone(), mask)                      # [bs, T, D]
        out["ctx" + tag] = ctx.reshape(Bi, Bc, T, D).numpy()
    # one masked variant: the last R // 3 keys of every pair masked (the -inf convention)
    maskk = torch.zeros(Bi * Bc, R, dtype=torch.bool); maskk[:, R - R // 3:] = True
    q = words.double().unsqueeze(0).expand(Bi, Bc, D, T).reshape(Bi * Bc, D, T)
    k = regions.double().unsqueeze(1).expand(Bi, Bc, D, R).reshape(Bi * Bc, D, R)
    out["ctx64_masked"] = ref(q.clone(), k.clone(), maskk).reshape(Bi, Bc, T, D).numpy()
    np.savez_compressed(os.path.join(HERE, f"ref_attn_{name}.npz"), words=words.numpy(), regions=regions.numpy(),
                        masked_from=R - R // 3, **out)
    print(f"ref_attn_{name}: |ctx|={float(np.linalg.norm(out['ctx64'])):.12f}")


if __name__ == "__main__":
    assert LR.reference_available(), "needs /root/reference"
    torch.set_num_threads(1)
    sim_case("sent_b32_d256_id", "sent", 32, 256, False, 0.5, 1)
    sim_case("sent_b24_d64_soft05", "sent", 24, 64, True, 0.5, 2)
    sim_case("sent_b24_d64_soft0", "sent", 24, 64, True, 0.0, 3)
    sim_case("sent_b88_d256_id", "sent", 88, 256, False, 0.0, 4, need=(True, False))
    sim_case("img_b16_d512_id", "img", 16, 512, False, 0.5, 5, need=(False, True))
    sim_case("img_b40_d512_soft05", "img", 40, 512, True, 0.5, 6, need=(False, True))
    magp_case("b6_3x8x8_d16", 6, (3, 8, 8), 16, 21, 0.08)
    magp_case("b5_3x7x9_d10", 5, (3, 7, 9), 10, 22, 0.11)        # row lengths not multiples of 4
    attn_case("i3_c4_d64_t6_r20", 3, 4, 64, 6, 20, 31)
    attn_case("i2_c3_d256_t18_r289", 2, 3, 256, 18, 289, 32)
    word_case("b6_d64_t7_r20", 6, 64, 7, 20, 11)
    word_case("b4_d256_t18_r289", 4, 256, 18, 289, 12, lean=0.15)
    word_case("b6_d128_t12_r64_nv", 6, 128, 12, 64, 13, normalize_values=True)
    word_case("b10_d64_t9_r33_soft", 10, 64, 9, 33, 14, b_global=True, smooth=0.5)
