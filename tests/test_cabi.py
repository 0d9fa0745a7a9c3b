"""The C-ABI shared library builds, loads, and exports every symbol include/mcgm.h declares
(no compute calls — this runs without a GPU)."""
import ctypes
import os
import re

from util import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcgm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcgm_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from mcgm_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mcgm.h but not exported"
    # the ctypes table binds exactly the declared surface
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_error_string():
    from mcgm_b200 import _lib
    lib = _lib.load()
    assert lib.mcgm_version() >= 100
    # argument validation happens before any CUDA call, so it is testable on the CPU box
    rc = lib.mcgm_pack_codebook(None, 0, 0, None, None)
    assert rc == -1
    assert b"pack_codebook" in lib.mcgm_last_error()
