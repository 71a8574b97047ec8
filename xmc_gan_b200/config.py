"""The one configuration knob the loss path reads.

The reference keeps a module-global EasyDict ``cfg`` (``xmc_gan/config/gan.py:7-90``) and its
losses read ``cfg.TRAIN.SMOOTH.GLOBAL`` at call time (``xmc_gan/train_gan.py:80,96,120``; default
0.5 at ``config/gan.py:41``, every shipped YAML sets ``0.``).  This mirror keeps the same access
path so a caller can do ``cfg.TRAIN.SMOOTH.GLOBAL = 0.`` exactly as with the reference.
"""
from types import SimpleNamespace

cfg = SimpleNamespace(
    TRAIN=SimpleNamespace(
        SMOOTH=SimpleNamespace(GLOBAL=0.5, MISMATCH=1.0, SENT=1.0, DISC=1.0),
        ENCODER_LOSS=SimpleNamespace(B_GLOBAL=False, SENT=False, WORD=False, DISC=False, VGG=False),
    )
)
