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

HOOKS_LIB_PATH = os.path.join(_HERE, "libxmcloss_hooks.so")   # tests / perf experiments only (-DXMC_TEST_HOOKS)
OBJ_DIR = os.path.join(_HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# translation units that contain `#ifdef XMC_TEST_HOOKS` code (debug dumps, pipeline traces, A/B switches)
HOOKED_SOURCES = ("wordregion_tc.cu", "prep.cu", "wordregion_split.cu", "region_head.cu")

XMC_F32, XMC_BF16 = 0, 1
PATH_FP32_SIMT, PATH_BF16_TCGEN05, PATH_FP32_TCGEN05 = 0, 1, 2

_vp, _i, _f, _sz, _ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong

# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "xmc_version": (_i, []),
    "xmc_last_error": (C.c_char_p, []),
    "xmc_check_device": (_i, []),
    "xmc_cosine_scores": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "xmc_cosine_scores_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xmc_infonce_combine_loss": (_i, [_vp, _i, _i, _i, _vp, _f, _i, _vp, _vp, _vp]),
    "xmc_simloss_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xmc_infonce_stats": (_i, [_vp, _i, _i, _vp, _i, _f, _vp, _vp, _vp]),
    "xmc_infonce_loss": (_i, [_vp, _vp, _i, _i, _vp, _vp, _f, _i, _i, _i, _i, _vp, _vp, _vp]),
    "xmc_infonce_grad": (_i, [_vp, _i, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _f, _i, _i, _vp, _vp, _vp]),
    "xmc_simloss_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _f, _vp, _vp,
                                  _vp, _vp, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xmc_simloss_workspace_bytes": (_sz, [_i, _i, _i]),
    "xmc_make_labels": (_i, [_vp, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "xmc_word_rows_compact": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "xmc_normalize_transpose": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "xmc_normalize_transpose_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "xmc_normalize_rows": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "xmc_normalize_rows_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "xmc_region_head_forward": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "xmc_region_head_backward_input": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "xmc_region_head_backward_weight": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "xmc_avgpool_rows": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "xmc_avgpool_rows_backward": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp]),
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


def _deps() -> list[str]:
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [HEADER]


def _stale(path: str = LIB_PATH) -> bool:
    if not os.path.isfile(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(nvcc, src, obj, extra, verbose):
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return res.stderr


def build(force: bool = False, verbose: bool = False, hooks: bool = True) -> str:
    """Compile libxmcloss.so for sm_100a with nvcc (no GPU needed).  Returns the .so path.

    One object per translation unit, compiled in parallel, linked twice: the product library and
    (``hooks``) ``libxmcloss_hooks.so``, the same code with ``-DXMC_TEST_HOOKS`` — the debug dumps,
    pipeline traces and A/B switches that tests and ``profiles/exp_*.py`` use.  The product library
    has none of them: it exports exactly what ``include/xmc_loss.h`` declares and holds no state.
    """
    want_hooks = hooks and (force or _stale(HOOKS_LIB_PATH))
    if not force and not _stale() and not want_hooks:
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []          # (src, obj, extra flags)
    plain, hooked = [], []
    for src in sources():
        base = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OBJ_DIR, base + ".o")
        jobs.append((src, obj, []))
        plain.append(obj)
        if hooks and os.path.basename(src) in HOOKED_SOURCES:
            hobj = os.path.join(OBJ_DIR, base + ".hooks.o")
            jobs.append((src, hobj, ["-DXMC_TEST_HOOKS"]))
            hooked.append(hobj)
        else:
            hooked.append(obj)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(lambda j: _compile(nvcc, j[0], j[1], j[2], verbose), jobs))
    if verbose:
        print("\n".join(logs))
    for out, objs in ((LIB_PATH, plain),) + (((HOOKS_LIB_PATH, hooked),) if hooks else ()):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", out, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
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


_hooks_lib = None


def hooks_lib() -> C.CDLL:
    """The -DXMC_TEST_HOOKS build (tests and perf experiments only; the product never loads it)."""
    global _hooks_lib
    with _lock:
        if _hooks_lib is None:
            if not os.path.isfile(HOOKS_LIB_PATH):
                raise RuntimeError(f"{HOOKS_LIB_PATH} is missing: build it with _lib.build(hooks=True)")
            h = C.CDLL(HOOKS_LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(h, name)
                fn.restype, fn.argtypes = res, args
            h.xmc_internal_set_debug_dump.argtypes = [_i]
            h.xmc_internal_set_prep_generic.argtypes = [_i]
            _hooks_lib = h
    return _hooks_lib


def check(rc: int, L=None) -> None:
    if rc != 0:
        msg = (L or lib()).xmc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libxmcloss error {rc}: {msg}")
