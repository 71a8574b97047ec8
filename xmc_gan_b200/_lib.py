"""ctypes binding of ``libxmcloss.so`` (C ABI declared in ``include/xmc_loss.h``).

The library is built in-tree by :func:`build` (plain ``nvcc`` for sm_100a; it cross-compiles
without a GPU) and loaded lazily by :func:`lib`.  There is no fallback of any kind: if the shared
object is missing or the device is not sm_100, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import re
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(ROOT, "include", "xmc_loss.h")
LIB_PATH = os.path.join(_HERE, "libxmcloss.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]

XMC_F32, XMC_BF16 = 0, 1
PATH_FP32_SIMT, PATH_BF16_TCGEN05 = 0, 1

_vp, _i, _f, _sz, _ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong

# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "xmc_version": (_i, []),
    "xmc_last_error": (C.c_char_p, []),
    "xmc_check_device": (_i, []),
    "xmc_cosine_scores": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "xmc_simloss_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xmc_infonce_stats": (_i, [_vp, _i, _i, _vp, _i, _f, _vp, _vp, _vp]),
    "xmc_infonce_loss": (_i, [_vp, _vp, _i, _i, _vp, _vp, _f, _i, _i, _i, _i, _vp, _vp]),
    "xmc_infonce_grad": (_i, [_vp, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _vp]),
    "xmc_simloss_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp,
                                  _vp, _vp, _f, _i, _i, _vp, _vp, _vp, _vp]),
    "xmc_make_labels": (_i, [_vp, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "xmc_word_rows_compact": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "xmc_normalize_transpose": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "xmc_normalize_transpose_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "xmc_wordregion_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "xmc_wordregion_forward": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xmc_wordregion_backward": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xmc_word_scores": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "xmc_word_scores_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "xmc_gradnorm_penalty_forward": (_i, [_vp, _ll, _vp, _ll, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "xmc_gradnorm_penalty_backward": (_i, [_vp, _ll, _vp, _ll, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _vp]),
    "xmc_infonce_combine_stats": (_i, [_vp, _i, _i, _vp, _vp]),
    "xmc_word_scores_infonce_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _i, _f, _vp, _vp, _vp, _vp, _f,
                                              _i, _i, _vp, _vp, _vp]),
}


def header_symbols() -> list[str]:
    """Every function name declared in include/xmc_loss.h."""
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(xmc_[a-z0-9_]+)\s*\(", text)))


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libxmcloss.so for sm_100a with nvcc (no GPU needed).  Returns the .so path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH, *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lock = threading.Lock()
_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the bound library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                    "xmc_gan_b200 has no CPU or PyTorch fallback.")
            h = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(h, name)
                fn.restype, fn.argtypes = res, args
            _lib = h
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().xmc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libxmcloss error {rc}: {msg}")
