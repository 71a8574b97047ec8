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


ATTENTION_FILE = "/root/reference/xmc_gan/model/concept_gan.py"


def load_reference_attention():
    """The reference's attention block ``CondConceptSampler.get_context_embs`` (``xmc_gan/model/concept_gan.py:532-555``)
    as a plain callable ``(queries [bs, D, Tq], keys [bs, D, Tk], mask [bs, Tk] bool) -> contexts [bs, Tq, D]``: cosine
    similarities of queries and keys (``F.normalize`` over the feature axis, ``matmul``), ``-inf`` on masked keys, softmax
    over the keys, sum of the UNIT keys.  The method's ``FunctionDef`` is lifted out of its class with ``ast`` and run
    with ``self = None`` (it uses no attribute) — unmodified, nothing copied.  The method averages the contexts over its
    query axis (``:553``), so it is called with ONE query per batch item and one group (``[bs*Tq, 1, D, 1]`` against
    ``[bs*Tq, 1, D, Tk]``), which makes that mean the identity.  (The sibling ``OutConceptBlock.get_context_embs``
    ``:374-394`` has the same body but multiplies ``[bs, p', C] x [bs, p', T]`` without transposing and only runs when
    the state count equals the feature width.)"""
    with open(ATTENTION_FILE, "r") as f:
        tree = ast.parse(f.read(), ATTENTION_FILE)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CondConceptSampler")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "get_context_embs")
    ns = {"torch": torch, "nn": torch.nn, "F": F}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), ATTENTION_FILE, "exec"), ns)

    def attention(queries, keys, mask):
        bs, D, Tq = queries.shape
        Tk = keys.shape[2]
        q = queries.transpose(1, 2).reshape(bs * Tq, 1, D, 1)                       # one query per item, one group
        k = keys.unsqueeze(1).expand(bs, Tq, D, Tk).reshape(bs * Tq, 1, D, Tk).clone()
        m = mask.unsqueeze(1).expand(bs, Tq, Tk).reshape(bs * Tq, Tk)
        return ns["get_context_embs"](None, q, k, m).view(bs, Tq, D)
    return attention
