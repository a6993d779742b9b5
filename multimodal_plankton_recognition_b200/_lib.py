"""Build + load ``libplk.so`` (the C ABI declared in ``include/plk.h``) through ctypes.

The library is built IN-TREE next to this file so that it travels with the repository
snapshot to the GPU box.  There is no fallback: if the library is missing and cannot be
built, or a call returns a non-zero status, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PLK_TRACE=1 (development only): a second library with in-kernel clock stamps (tools/trace_tc.py)
_TRACE = os.environ.get("PLK_TRACE", "0") == "1"
# PLK_VARIANT_FLAGS="-DX=1 ..." (development only): a third library built with extra nvcc flags (kernel A/B runs)
_VARIANT = os.environ.get("PLK_VARIANT_FLAGS", "").split()
LIB_PATH = os.path.join(HERE, "libplk_trace.so" if _TRACE else ("libplk_variant.so" if _VARIANT else "libplk.so"))
OBJ_DIR = os.path.join(HERE, "build_trace" if _TRACE else ("build_variant" if _VARIANT else "build"))
SOURCES = ["api.cu", "elementwise.cu", "infonce_simt.cu", "topk_simt.cu", "tc_host.cu", "stager.cu",
           "infonce_tc.cu", "topk_tc.cu", "proj_tc.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"] + (["-DPLK_TRACE"] if _TRACE else []) + _VARIANT

PLK_F32, PLK_BF16, PLK_F16 = 0, 1, 2

_lib = None


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libplk.so cannot be built (no CPU fallback exists)")
    return exe


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "plk.h")]
    return any(os.path.getmtime(p) > t for p in deps if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link ``libplk.so`` (no GPU needed)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(one, SOURCES))
    tmp = LIB_PATH + ".tmp"
    r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_i64, _int, _vp, _sz = C.c_int64, C.c_int, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/plk.h one to one
SIGNATURES = {
    "plk_version": (_int, []),
    "plk_last_error": (C.c_char_p, []),
    "plk_device_supports_tc": (_int, []),
    "plk_launch_count": (_i64, []),
    "plk_l2norm_fwd": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _vp, _vp, _vp, _int, _vp]),
    "plk_l2norm_pair_fwd": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _int, _i64, _vp, _vp, _vp, _vp, _vp, _i64,
                                   _vp, _i64, _vp]),
    "plk_infonce_fwd": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _int, _vp]),
    "plk_infonce_grad_finish_pair": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                            _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plk_infonce_loss_xgpu": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _vp,
                                     _vp]),
    "plk_infonce_loss": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "plk_infonce_grad_parts": (_int, [_int, _i64, _i64, _i64, _i64]),
    "plk_infonce_grad_pair_parts": (_int, [_int, _i64, _i64, _i64, _i64]),
    "plk_infonce_grad_pair": (_int, [_vp, _vp, _vp, _vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp]),
    "plk_infonce_grad": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plk_infonce_grad_finish": (_int, [_vp, _int, _vp, _vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _vp, _i64, _vp, _int, _vp]),
    "plk_infonce_grad_finish_pair_xgpu": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                 _int, _int, _vp, _vp, _vp]),
    "plk_infonce_dls": (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "plk_clip_loss_state_bytes": (_sz, [_int, _i64, _i64]),
    "plk_clip_loss_workspace_bytes": (_sz, [_int, _i64, _i64, _i64]),
    "plk_clip_loss_forward": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _vp]),
    "plk_clip_loss_forward_xgpu": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int,
                                          _int, _vp, _vp, _vp]),
    "plk_clip_loss_backward": (_int, [_vp, _vp, C.c_float, _vp, _vp, _i64, _i64, _i64, _int, _i64, _i64, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _vp]),
    "plk_clip_loss_backward_xgpu": (_int, [_vp, _vp, C.c_float, _vp, _vp, _i64, _i64, _i64, _int, _i64, _i64, _vp,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp]),
    "plk_siglip_loss_forward": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    "plk_siglip_loss_backward": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp]),
    "plk_siglip_loss": (_int, [_vp, _i64, _vp, _vp]),
    "plk_siglip_grad_finish_pair": (_int, [_vp, _vp, _int, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "plk_project_normalise": (_int, [_vp, _i64, _vp, _i64, _int, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "plk_stager_create": (_vp, [_int, _int]),
    "plk_stager_destroy": (None, [_vp]),
    "plk_stager_issue": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _vp, _sz]),
    "plk_stager_acquire": (_int, [_vp, _int, _int, _vp]),
    "plk_stager_read_async": (_int, [_vp, _int, _vp, _vp, _sz, _vp]),
    "plk_stager_read_wait": (_int, [_vp, _int]),
    "plk_topk_workspace_bytes": (_sz, [_i64, _i64, _i64, _int, _int]),
    "plk_topk_candidates": (_int, [_vp, _vp, _int, _i64, _vp, _i64, _i64, _i64, _int, _i64, _vp, _vp, _vp,
                                   _sz, _vp]),
    "plk_topk_rescore": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _i64, _int, _vp, _vp, _vp, _vp]),
    "plk_topk_merge": (_int, [_vp, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "plk_knn_vote": (_int, [_vp, _vp, _i64, _int, _vp, _i64, _vp, _vp]),
}


class PlkError(RuntimeError):
    pass


class _Lib:
    def __init__(self, path):
        self.path = path
        self.cdll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.cdll, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc, what=""):
        if rc != 0:
            msg = self.plk_last_error().decode("utf-8", "replace")
            raise PlkError(f"libplk {what} failed (status {rc}): {msg}")


def load(build_if_missing: bool = True) -> _Lib:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run __graft_entry__.build() (no CPU fallback exists)")
        build()
    _lib = _Lib(LIB_PATH)
    return _lib
