"""CPU oracle for the XMC-GAN contrastive-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``xmc_gan_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Parity status
-------------
* ``cosine_scores`` / ``make_labels`` / ``sent_loss`` / ``img_loss``:
  restatements of ``xmc_gan/train_gan.py:72-139``.  PINNED: checked against the
  reference's own source (AST-loaded, unmodified, executed on CPU) by
  ``tests/golden/make_golden.py`` in the build container; the resulting
  input/output vectors are committed under ``tests/golden/`` and re-checked by
  ``tests/test_oracle.py`` everywhere (the GPU box has no ``/root/reference``).
* ``magp_penalty``: restatement of the MA-GP reduction ``xmc_gan/train_gan.py:244-249`` (inline
  statements of ``train()``).  PINNED the same way: ``load_reference.load_reference_magp`` lifts those six
  assignments out of the reference's ``train`` with ``ast`` and runs them; vectors in
  ``tests/golden/ref_magp_*.npz``.
* ``word_scores`` / ``word_loss``: PARITY UNPINNED.  The reference names
  ``word_loss`` (``xmc_gan/train_gan.py:220-222, 267-269``) but only raises
  ``NotImplementedError``; the restatement here follows the XMC-GAN paper's
  word-region formulation and the reference's surrounding conventions (see
  ``oracle/word_region.py``).  Stage pins: ``word_region.attend`` (cosines, softmax over the keys, contexts) against the
  reference's attention block ``xmc_gan/model/concept_gan.py:532-555`` (``tests/golden/ref_attn_*.npz``), the InfoNCE tail
  against ``sent_loss``.
"""
from .ref_losses import (  # noqa: F401
    cosine_scores, make_labels, infonce_tail, sent_loss, img_loss, num_pos_of, magp_penalty,
)
from .word_region import attend, word_scores, word_loss  # noqa: F401
