"""GPU parity of the SHARDED (global-negatives) path: W ranks as W processes on ONE GPU.

Every rank runs the product code — ``xmc_gan_b200.losses`` with the CUDA backend — on its rows of the global
problem: rectangular score matrices (``Bk = W * Bq``), non-zero ``diag_offset``, column ranges that start
past 0, the packet exchange and the gradient reduce-scatter.  The collectives travel over gloo with the
tensors staged through the host (NCCL refuses two ranks on one device); the kernels and the host logic are
the product's.  Definition of correct: each rank's loss equals the single-process oracle on the
concatenated global batch and its gradients equal the oracle's gradient rows of that rank.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
RHO = (4.0, 5.0, 6.0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(seed, Bg, D, Dimg, Dw, T_, R):
    g = torch.Generator().manual_seed(seed)
    words = torch.randn(Bg, Dw, T_, generator=g)
    regions = torch.randn(Bg, Dw, R, generator=g) + 0.3 * words[:, :, torch.randint(0, T_, (R,), generator=g)]
    d = dict(img=torch.randn(Bg, D, generator=g), sent=torch.randn(Bg, D, generator=g),
             real=torch.randn(Bg, Dimg, generator=g), fake=torch.randn(Bg, Dimg, generator=g),
             words=words, regions=regions)
    # soft positives inside a rank and across ranks (make_labels, train_gan.py:76-82)
    d["sent"][3] = d["sent"][Bg - 2] + 0.05 * torch.randn(D, generator=g)
    d["sent"][1] = d["sent"][2] + 0.05 * torch.randn(D, generator=g)
    lens = torch.randint(1, T_ + 1, (Bg,), generator=g)
    lens[Bg // 2] = 0                                           # one caption that is all padding
    d["mask"] = torch.arange(T_).unsqueeze(0) >= lens.unsqueeze(1)
    return d


CASES = [  # name, b_global, SMOOTH.GLOBAL, precision, fused
    ("identity-fp32", False, 0.5, "fp32", False),
    ("identity-bf16", False, 0.5, "bf16", False),
    ("identity-bf16-fused", False, 0.5, "bf16", True),
    ("soft0.5-bf16-fused", True, 0.5, "bf16", True),
    ("soft0-fp32-fused", True, 0.0, "fp32", True),
    ("soft0.5-fp32", True, 0.5, "fp32", False),
]


def _shape(world):
    return dict(B=24 if world == 2 else 8, D=256, Dimg=512, Dw=256, T_=9, R=70)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), XMC_CHECK_ERRORS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from xmc_gan_b200 import losses as L
        from xmc_gan_b200 import train_gan as T

        class HostStagedComm(L.Comm):
            """The product's Comm with its two primitives carried by gloo through host memory."""

            def __init__(self, group):
                super().__init__(group)
                self.coalesce = False

            def _all_gather(self, outs, ins):
                for o, i in zip(outs, ins):
                    h = i.detach().contiguous().cpu().view(torch.uint8)
                    oh = torch.empty((self.world,) + tuple(h.shape), dtype=torch.uint8)
                    dist.all_gather_into_tensor(oh, h.unsqueeze(0), group=self.group)
                    o.copy_(oh.view(-1).view(o.dtype).view(o.shape))
                return L._Done()

            def _reduce_scatter(self, outs, ins):
                for o, i in zip(outs, ins):
                    h = i.detach().float().cpu().contiguous()
                    oh = torch.empty((h.shape[0] // self.world,) + tuple(h.shape[1:]), dtype=torch.float32)
                    dist.reduce_scatter_tensor(oh, h, op=dist.ReduceOp.SUM, group=self.group)
                    o.copy_(oh)
                return L._Done()

        comm = HostStagedComm(dist.group.WORLD)
        sh = _shape(world)
        B = sh["B"]
        d = _data(7, B * world, sh["D"], sh["Dimg"], sh["Dw"], sh["T_"], sh["R"])
        sl = slice(rank * B, (rank + 1) * B)
        res = {}
        for name, b_global, smooth, precision, fused in CASES:
            T.cfg.TRAIN.SMOOTH.GLOBAL = smooth
            dt = torch.bfloat16 if precision == "bf16" else torch.float32
            leaf = lambda x: x[sl].to(dt).cuda().requires_grad_()
            img, sent, fake, words, regions = leaf(d["img"]), leaf(d["sent"]), leaf(d["fake"]), leaf(d["words"]), leaf(d["regions"])
            real = d["real"][sl].to(dt).cuda()
            mask = d["mask"][sl].cuda()
            labels = T.make_labels(B, d["sent"][sl].cuda(), b_global, group=comm)
            if fused:
                parts = T.contrastive_losses(img, sent, real, fake, regions, words, mask, labels, b_global,
                                             rho1=RHO[0], rho2=RHO[1], rho3=RHO[2], precision=precision, group=comm)
            else:
                parts = (T.sent_loss(img, sent, labels, b_global, group=comm),
                         T.img_loss(real, fake, labels, b_global, group=comm),
                         T.word_loss(regions, words, mask, labels, b_global, rho1=RHO[0], rho2=RHO[1], rho3=RHO[2],
                                     precision=precision, group=comm))
            (parts[0] + 0.5 * parts[1] + 2.0 * parts[2]).backward()
            torch.cuda.synchronize()
            res[name] = dict(parts=[float(p.detach()) for p in parts], labels=labels.detach().cpu(),
                             grads=[t.grad.float().cpu() for t in (img, sent, fake, words, regions)])
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_cuda_path_matches_oracle_on_the_concatenated_batch(world, tmp_path):
    import oracle
    from util import TOL_BF16, TOL_FP32
    # results come back through files: a multiprocessing.Manager would fork this (CUDA-initialised) process
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = {r: torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)}
    sh = _shape(world)
    B, Bg = sh["B"], sh["B"] * world
    d = _data(7, Bg, sh["D"], sh["Dimg"], sh["Dw"], sh["T_"], sh["R"])
    for name, b_global, smooth, precision, fused in CASES:
        tol = TOL_BF16 if precision == "bf16" else TOL_FP32
        rd = (lambda x: x.bfloat16().double()) if precision == "bf16" else (lambda x: x.double())   # the oracle sees what the kernels see
        leaf = lambda x: rd(x).clone().requires_grad_()
        img, sent, fake, words, regions = leaf(d["img"]), leaf(d["sent"]), leaf(d["fake"]), leaf(d["words"]), leaf(d["regions"])
        labels = oracle.make_labels(Bg, d["sent"], b_global, smooth_global=smooth)
        parts = [oracle.sent_loss(img, sent, labels, b_global, smooth),
                 oracle.img_loss(rd(d["real"]), fake, labels, b_global, smooth),
                 oracle.word_loss(regions, words, d["mask"], labels, b_global, smooth, *RHO)]
        (parts[0] + 0.5 * parts[1] + 2.0 * parts[2]).backward()
        if b_global:
            assert (labels - torch.eye(Bg)).abs().sum() > 0, "test data must contain soft positives"
        refs = (img.grad, sent.grad, fake.grad, words.grad, regions.grad)
        for rank in range(world):
            sl = slice(rank * B, (rank + 1) * B)
            r = out[rank][name]
            assert torch.equal(r["labels"], labels[sl]), (name, rank, "label rows")
            for k, (got, ref) in enumerate(zip(r["parts"], parts)):
                rel = abs(got - float(ref.detach())) / abs(float(ref.detach()))
                assert rel <= tol, (name, rank, "loss", k, got, float(ref.detach()))
            for k, (got, ref) in enumerate(zip(r["grads"], refs)):
                err = float((got.double() - ref[sl]).norm() / ref[sl].norm())
                assert err <= tol, (name, rank, "grad", k, err)
