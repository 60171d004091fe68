"""The C-ABI shared library loads on a GPU-less box and exports every symbol that
include/evo_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from evo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "evo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(evo_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m evo_b200.build` (or __graft_entry__.build())"
    assert os.path.dirname(_lib.LIB_PATH) == os.path.join(ROOT, "evo_b200")


def test_every_declared_symbol_is_exported():
    names = header_functions()
    assert len(names) >= 18
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"declared in include/evo_b200.h but not exported: {missing}"


def test_binding_table_covers_the_header():
    assert sorted(_lib.SIGNATURES) == header_functions()


def test_loader_binds_and_reports_version():
    lib = _lib.lib()
    assert lib.evo_abi_version() == 1
    assert lib.evo_last_error() is not None
    lib.evo_reset_launch_count()
    assert lib.evo_launch_count() == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.EvoError, match="no CPU"):
        _lib.lib()


def test_sass_has_tcgen05_and_tma():
    """Blackwell-native evidence: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "HMMA.16816" not in sass     # no legacy mma.sync tensor path
