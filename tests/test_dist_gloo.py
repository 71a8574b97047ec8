"""world_size-2 gloo tests (CPU) of the multi-rank host logic: all-gather of the column operand,
rank-offset identity labels, column-statistics exchange, reduce-scatter of gradients, global
soft-positive labels.  The compute kernels are replaced by the CPU checker backend
(tests/cpu_ops.py) — the product backend needs a GPU; its own parity is tested with -m gpu.

Definition of correct: each rank's loss equals the single-process oracle on the
concatenated global batch, and its gradients equal the oracle's gradient rows of that rank."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(seed, Bg, D, Dw, T_, R):
    g = torch.Generator().manual_seed(seed)
    d = dict(
        img=torch.randn(Bg, D, generator=g, dtype=torch.float64),
        sent=torch.randn(Bg, D, generator=g, dtype=torch.float64),
        words=torch.randn(Bg, Dw, T_, generator=g, dtype=torch.float64),
        regions=torch.randn(Bg, Dw, R, generator=g, dtype=torch.float64),
    )
    d["sent"][3] = d["sent"][Bg - 2] + 0.05 * torch.randn(D, generator=g, dtype=torch.float64)   # cross-rank positives
    d["sent"][1] = d["sent"][2] + 0.05 * torch.randn(D, generator=g, dtype=torch.float64)
    lens = torch.randint(1, T_ + 1, (Bg,), generator=g)
    d["mask"] = torch.arange(T_).unsqueeze(0) >= lens.unsqueeze(1)
    return d


def _worker(rank, port, b_global, smooth, precision, Dw, out, fused=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from cpu_ops import CpuOps
        from xmc_gan_b200 import train_gan as T
        torch.set_num_threads(1)
        T.cfg.TRAIN.SMOOTH.GLOBAL = smooth
        ops = CpuOps()
        B, D, T_, R = 6, 16, 5, 7
        d = _data(0, B * WORLD, D, Dw, T_, R)
        sl = slice(rank * B, (rank + 1) * B)
        leaf = lambda x: x[sl].clone().requires_grad_()
        img, sent, words, regions = leaf(d["img"]), leaf(d["sent"]), leaf(d["words"]), leaf(d["regions"])
        group = dist.group.WORLD
        labels = T.make_labels(B, d["sent"][sl].float(), b_global, group=group, _ops=ops)
        if fused:      # the three losses through ONE autograd function: grouped gather / exchange / reduce-scatter
            real, fake = leaf(d["img"]).detach(), leaf(d["sent"])
            l3 = T.contrastive_losses(img, sent, real, fake, regions, words, d["mask"][sl], labels, b_global,
                                      rho1=4.0, rho2=5.0, rho3=6.0, precision=precision, group=group, _ops=ops)
            loss = l3[0] + 0.5 * l3[1] + l3[2]
            loss.backward()
            out[rank] = dict(loss=loss.detach(), labels=labels.detach().clone(), parts=[x.detach() for x in l3],
                             grads=[t.grad.clone() for t in (img, sent, words, regions, fake)])
            return
        loss = (T.sent_loss(img, sent, labels, b_global, group=group, _ops=ops)
                + T.img_loss(sent.detach(), img, labels, b_global, tau=0.5, group=group, _ops=ops)
                + T.word_loss(regions, words, d["mask"][sl], labels, b_global, rho1=4.0, rho2=5.0, rho3=6.0,
                              precision=precision, group=group, _ops=ops))
        loss.backward()
        assert ops.compactions == 1          # padding words are compacted away on both paths
        out[rank] = dict(loss=loss.detach(), labels=labels.detach().clone(),
                         grads=[t.grad.clone() for t in (img, sent, words, regions)])
    finally:
        dist.destroy_process_group()


# precision "bf16" selects the tcgen05-path host flow (saved contexts, pre-filled accumulators); both paths
# compact the padding words away (device-side row count, gradients scattered back through row_of).  The
# checker backend computes in fp64 either way.
@pytest.mark.parametrize("b_global,smooth,precision,Dw", [(False, 0.5, None, 8), (True, 0.5, None, 8), (True, 0.0, None, 8),
                                                          (True, 0.5, "bf16", 128), (False, 0.5, "bf16", 128)])
def test_two_rank_global_negatives_match_single_process_oracle(b_global, smooth, precision, Dw):
    import oracle
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), b_global, smooth, precision, Dw, out), nprocs=WORLD, join=True)

    B, D, T_, R = 6, 16, 5, 7
    Bg = B * WORLD
    d = _data(0, Bg, D, Dw, T_, R)
    leaf = lambda x: x.clone().requires_grad_()
    img, sent, words, regions = leaf(d["img"]), leaf(d["sent"]), leaf(d["words"]), leaf(d["regions"])
    labels = oracle.make_labels(Bg, d["sent"].float(), b_global, smooth_global=smooth)
    num_pos = oracle.num_pos_of(labels, b_global, smooth)
    loss = (oracle.sent_loss(img, sent, labels, b_global, smooth)
            + oracle.infonce_tail(oracle.cosine_scores(sent.detach(), img) / 0.5, labels, num_pos)
            + oracle.word_loss(regions, words, d["mask"], labels, b_global, smooth, 4.0, 5.0, 6.0))
    loss.backward()
    if b_global:
        assert (labels - torch.eye(Bg)).abs().sum() > 0, "test data must contain soft positives"
    for rank in range(WORLD):
        sl = slice(rank * B, (rank + 1) * B)
        r = out[rank]
        assert torch.equal(r["labels"], labels[sl]), f"rank {rank} label rows"
        assert abs(float(r["loss"]) - float(loss.detach())) < 1e-6 * abs(float(loss.detach())), (float(r["loss"]), float(loss.detach()))
        for got, ref in zip(r["grads"], (img.grad, sent.grad, words.grad, regions.grad)):
            err = float((got - ref[sl]).norm() / ref[sl].norm())
            assert err < 1e-6, (rank, err)


@pytest.mark.parametrize("b_global,smooth,precision,Dw", [(False, 0.5, None, 8), (True, 0.5, "bf16", 128), (True, 0.0, None, 8)])
def test_two_rank_fused_losses_match_single_process_oracle(b_global, smooth, precision, Dw):
    """contrastive_losses (one grouped all-gather, one packet exchange, one grouped reduce-scatter per step)
    against the single-process oracle on the concatenated batch, every loss and every gradient."""
    import oracle
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), b_global, smooth, precision, Dw, out, True), nprocs=WORLD, join=True)

    B, D, T_, R = 6, 16, 5, 7
    Bg = B * WORLD
    d = _data(0, Bg, D, Dw, T_, R)
    leaf = lambda x: x.clone().requires_grad_()
    img, sent, words, regions, fake = leaf(d["img"]), leaf(d["sent"]), leaf(d["words"]), leaf(d["regions"]), leaf(d["sent"])
    labels = oracle.make_labels(Bg, d["sent"].float(), b_global, smooth_global=smooth)
    parts = [oracle.sent_loss(img, sent, labels, b_global, smooth),
             oracle.img_loss(d["img"], fake, labels, b_global, smooth),
             oracle.word_loss(regions, words, d["mask"], labels, b_global, smooth, 4.0, 5.0, 6.0)]
    loss = parts[0] + 0.5 * parts[1] + parts[2]
    loss.backward()
    for rank in range(WORLD):
        sl = slice(rank * B, (rank + 1) * B)
        r = out[rank]
        for got, ref in zip(r["parts"], parts):
            assert abs(float(got) - float(ref.detach())) < 1e-6 * abs(float(ref.detach())), (rank, float(got), float(ref))
        for got, ref in zip(r["grads"], (img.grad, sent.grad, words.grad, regions.grad, fake.grad)):
            err = float((got - ref[sl]).norm() / ref[sl].norm())
            assert err < 1e-6, (rank, err)


def test_cosine_scores_is_differentiable_like_the_reference():
    """train_gan.cosine_scores carries a gradient (the reference's is an ordinary torch expression, :85-91)."""
    import oracle
    sys.path.insert(0, HERE)
    from cpu_ops import CpuOps
    from xmc_gan_b200 import train_gan as T
    g = torch.Generator().manual_seed(3)
    a = torch.randn(5, 12, generator=g, dtype=torch.float64).requires_grad_()
    b = torch.randn(7, 12, generator=g, dtype=torch.float64).requires_grad_()
    w = torch.randn(5, 7, generator=g, dtype=torch.float64)
    (T.cosine_scores(a, b, _ops=CpuOps()) * w).sum().backward()
    a2, b2 = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    (oracle.cosine_scores(a2, b2) * w).sum().backward()
    # the upstream gradient crosses the boundary as fp32 (the C ABI's type): 1e-6, not 1e-12
    assert torch.allclose(a.grad, a2.grad, atol=1e-6) and torch.allclose(b.grad, b2.grad, atol=1e-6)


def test_identity_tag_does_not_survive_an_edit():
    """make_labels tags identity labels so the kernels skip reading them; an in-place edit voids the tag."""
    from xmc_gan_b200 import losses as L
    lab = L.tag_identity(torch.eye(4))
    assert L._is_identity(lab, (4, 4))
    assert not L._is_identity(lab, (4, 8))
    lab[0, 1] = 0.5
    assert not L._is_identity(lab, (4, 4))


def test_channels_last_regions_take_the_row_layout_path():
    """SURVEY §8f N2 host logic: a channels-last region map is consumed as [B, R, D] rows (no transposing layout
    kernels) and its gradient comes back channels-last; same numbers as the contiguous map."""
    import oracle
    sys.path.insert(0, HERE)
    from cpu_ops import CpuOps
    from xmc_gan_b200 import train_gan as T

    class Counting(CpuOps):
        rows = 0

        def normalize_rows(self, *a, **k):
            self.rows += 1
            return super().normalize_rows(*a, **k)

    g = torch.Generator().manual_seed(4)
    B, D, H, W, T_ = 5, 8, 3, 4, 6
    reg = torch.randn(B, D, H, W, generator=g, dtype=torch.float64)
    words = torch.randn(B, D, T_, generator=g, dtype=torch.float64)
    mask = torch.arange(T_)[None] >= torch.randint(1, T_ + 1, (B, 1), generator=g)
    ops = Counting()
    r_cl = reg.clone().contiguous(memory_format=torch.channels_last).requires_grad_()
    loss = T.word_loss(r_cl, words, mask, T.make_labels(B, None, False, device=torch.device("cpu"), _ops=ops), False, _ops=ops)
    loss.backward()
    assert ops.rows == 1 and r_cl.grad.is_contiguous(memory_format=torch.channels_last)
    r_ref = reg.clone().requires_grad_()
    lo = oracle.word_loss(r_ref, words, mask, torch.eye(B), False)
    lo.backward()
    assert abs(float(loss) - float(lo)) < 1e-6 * abs(float(lo))
    assert float((r_cl.grad - r_ref.grad).norm() / r_ref.grad.norm()) < 1e-6
