"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every
symbol include/plk.h declares, with the argument counts the ctypes binding uses."""
import os
import re

import pytest

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "plk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|int64_t|const char\*|void\*|void)\s+(plk_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_library_builds_and_exports_every_declared_symbol():
    from multimodal_plankton_recognition_b200 import _lib
    _lib.build()
    lib = _lib.load()
    decl = _header_functions()
    assert len(decl) >= 16
    for name, nargs in decl.items():
        assert hasattr(lib.cdll, name), f"{name} declared in plk.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args"
    assert set(_lib.SIGNATURES) == set(decl)
    assert lib.plk_version() >= 100
    assert lib.plk_launch_count() >= 0


def test_argument_validation_happens_before_any_launch():
    """Bad arguments are rejected with a status + message, no CUDA call needed (runs without a GPU)."""
    from multimodal_plankton_recognition_b200 import _lib
    lib = _lib.load()
    rc = lib.plk_infonce_fwd(None, None, 0, 8, 4, 0, 4, 8, 4, None, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.plk_last_error()
    rc = lib.plk_topk_candidates(1, 1, 0, 8, 1, 4, 4, 8, 100, 0, 1, 1, None, 0, None)
    assert rc == -1 and b"kc" in lib.plk_last_error()
    with pytest.raises(_lib.PlkError, match="status -1"):
        lib.check(rc, "plk_topk_candidates")
    assert lib.plk_infonce_grad_parts(0, 4096, 4096, 256, 4096) == 1
    assert lib.plk_infonce_grad_parts(1, 4096, 4096, 256, 4096) == 4
