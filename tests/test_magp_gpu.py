"""Parity of the MA-GP reduction (xmc_gradnorm_penalty_forward/backward) with the reference's own statements
(golden vectors of train_gan.py:244-249) and with the oracle, including the double backward the penalty is
used for (the gradients fed in come from autograd.grad(..., create_graph=True))."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from util import TOL_FP32, lerr, nerr

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_magp_*.npz")))


@pytest.fixture(scope="module")
def T():
    from xmc_gan_b200 import train_gan
    return train_gan


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_reference_vectors(T, path):
    z = np.load(path)
    a = torch.from_numpy(z["g0"]).cuda().requires_grad_()
    b = torch.from_numpy(z["g1"]).cuda().requires_grad_()
    loss = T.magp_penalty((a, b))
    loss.backward()
    assert lerr(loss, z["loss64"]) <= TOL_FP32
    assert nerr(a.grad, torch.from_numpy(z["d0_64"])) <= TOL_FP32
    assert nerr(b.grad, torch.from_numpy(z["d1_64"])) <= TOL_FP32


@pytest.mark.parametrize("B,shape,D", [(8, (3, 32, 32), 256), (3, (3, 17, 5), 7), (16, (3, 64, 64), 256), (1, (1,), 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vs_oracle(T, B, shape, D, dtype):
    g = torch.Generator().manual_seed(B + D)
    gi = (torch.randn(B, *shape, generator=g) * 0.05).to(dtype)
    gs = (torch.randn(B, D, generator=g) * 0.05).to(dtype)
    a = gi.cuda().requires_grad_(); b = gs.cuda().requires_grad_()
    loss = T.magp_penalty((a, b))
    (loss * 0.7).backward()                                   # upstream gradient != 1
    ar = gi.double().requires_grad_(); br = gs.double().requires_grad_()   # oracle on the values the kernel saw
    ref = oracle.magp_penalty(ar, br)
    (ref * 0.7).backward()
    tol = TOL_FP32 if dtype == torch.float32 else 8e-3        # bf16: the returned gradients are rounded to bf16
    assert lerr(loss, ref) <= TOL_FP32
    assert nerr(a.grad, ar.grad) <= tol and nerr(b.grad, br.grad) <= tol
    assert a.grad.dtype == dtype and a.grad.shape == a.shape


def test_only_one_input_needs_grad_and_other_powers(T):
    g = torch.Generator().manual_seed(0)
    gi = torch.randn(4, 3, 8, 8, generator=g) * 0.2
    gs = torch.randn(4, 16, generator=g) * 0.2
    a = gi.cuda().requires_grad_(); b = gs.cuda()
    loss = T.magp_penalty((a, b), power=4.0, weight=0.5)
    loss.backward()
    ar = gi.double().requires_grad_()
    ref = oracle.magp_penalty(ar, gs.double(), power=4.0, weight=0.5)
    ref.backward()
    assert lerr(loss, ref) <= TOL_FP32 and nerr(a.grad, ar.grad) <= TOL_FP32 and b.grad is None


def test_double_backward_through_a_discriminator(T):
    """The use in train_gan.py:233-252: penalty of d out / d (image, sentence), back-propagated into the weights."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(3 * 6 * 6 + 10, 32), torch.nn.Softplus(), torch.nn.Linear(32, 1)).cuda()
    x = torch.randn(5, 3, 6, 6, device="cuda").requires_grad_()
    s = torch.randn(5, 10, device="cuda").requires_grad_()

    def penalty(fn):
        out = net(torch.cat((x.flatten(1), s), 1))
        grads = torch.autograd.grad(out, (x, s), grad_outputs=torch.ones_like(out), retain_graph=True, create_graph=True)
        return fn(grads)

    net.zero_grad()
    penalty(lambda gr: T.magp_penalty(gr)).backward()
    grab = lambda: [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    got = grab()                                   # the last bias does not influence d out / d input: no gradient
    net.zero_grad()
    penalty(lambda gr: oracle.magp_penalty(gr[0], gr[1])).backward()
    ref = grab()
    assert sum(g is not None for g in ref) >= 3
    for a, b in zip(got, ref):
        assert (a is None) == (b is None)
        if b is not None and float(b.norm()) > 0:
            assert nerr(a, b) <= TOL_FP32


def test_cpu_tensor_raises(T):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.magp_penalty((torch.randn(2, 3), torch.randn(2, 3)))
