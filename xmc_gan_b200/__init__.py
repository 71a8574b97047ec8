"""xmc_gan_b200 — B200-native (sm_100a) implementation of XMC-GAN's cross-modal contrastive-loss
path: sentence–image InfoNCE, word–region attention contrastive loss, real–fake image InfoNCE,
forward and backward, as drop-in ``torch.autograd`` ops with the reference's signatures.

    from xmc_gan_b200.train_gan import make_labels, cosine_scores, sent_loss, img_loss, word_loss

Host side: Python/PyTorch (device memory, streams, ``torch.distributed``).  Compute: hand-written
CUDA kernels in ``csrc/`` behind the C ABI of ``include/xmc_loss.h`` (``libxmcloss.so``, loaded
with ctypes).  No Triton, no multi-backend dispatch, no CPU fallback.
"""
from .config import cfg  # noqa: F401

__version__ = "0.1.0"
