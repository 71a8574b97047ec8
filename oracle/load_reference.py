"""Load the reference's own loss functions, unmodified, for pinning the oracle.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``import xmc_gan.train_gan``
is impossible in this image (``easydict``, ``sentence_transformers`` and
``pytorch_fid`` are absent and the module has import-time side effects), so the
four ``FunctionDef`` nodes ``make_labels``, ``cosine_scores``, ``sent_loss`` and
``img_loss`` (``xmc_gan/train_gan.py:72-139``) are parsed out of the read-only
file with ``ast`` and executed in a namespace holding ``torch``, ``F`` and a
stub ``cfg``.  No reference source is copied into this repo.

Only usable where ``/root/reference`` exists (the build container); the GPU box
has no such path, so nothing that runs there may call this.
"""
from __future__ import annotations

import ast
import contextlib
import os
from types import SimpleNamespace

import torch
import torch.nn.functional as F

REFERENCE_FILE = "/root/reference/xmc_gan/train_gan.py"
_WANTED = ("make_labels", "cosine_scores", "sent_loss", "img_loss")


def reference_available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


def load_reference_losses(smooth_global: float = 0.5) -> SimpleNamespace:
    """Return a namespace with the reference's four functions and its ``cfg`` stub."""
    with open(REFERENCE_FILE, "r") as f:
        tree = ast.parse(f.read(), REFERENCE_FILE)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in _WANTED]
    assert sorted(n.name for n in keep) == sorted(_WANTED), "reference layout changed"
    cfg = SimpleNamespace(TRAIN=SimpleNamespace(SMOOTH=SimpleNamespace(GLOBAL=smooth_global)))
    ns = {"torch": torch, "F": F, "cfg": cfg}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REFERENCE_FILE, "exec"), ns)
    return SimpleNamespace(cfg=cfg, **{k: ns[k] for k in _WANTED})


@contextlib.contextmanager
def cuda_is_identity():
    """``make_labels`` hard-codes ``.cuda()`` (train_gan.py:74); make it a no-op on CPU."""
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = saved


_MAGP_TARGETS = ("grad0", "grad1", "grad", "grad_l2norm", "d_loss_gp", "d_loss")


def load_reference_magp():
    """The reference's MA-GP reduction (``train_gan.py:244-249``) as a callable ``grads -> d_loss``.

    The six assignments are statements inside ``train()`` (under ``if cfg.TRAIN.MAGP:``), not a function:
    they are lifted out of the parsed file by their target names, in source order, and executed with
    ``grads`` bound — unmodified, nothing copied."""
    with open(REFERENCE_FILE, "r") as f:
        tree = ast.parse(f.read(), REFERENCE_FILE)
    train = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train")
    stmts = []
    for node in ast.walk(train):
        if isinstance(node, ast.If) and "MAGP" in ast.unparse(node.test):
            stmts = [st for st in node.body if isinstance(st, ast.Assign) and len(st.targets) == 1
                     and isinstance(st.targets[0], ast.Name) and st.targets[0].id in _MAGP_TARGETS]
            break
    assert [st.targets[0].id for st in stmts] == list(_MAGP_TARGETS), "reference layout changed"
    code = compile(ast.Module(body=stmts, type_ignores=[]), REFERENCE_FILE, "exec")

    def d_loss(grads):
        ns = {"torch": torch, "grads": grads}
        exec(code, ns)
        return ns["d_loss"]
    return d_loss
