"""CPU: the C-ABI library builds, loads and exports every symbol include/grl_b200.h declares."""
import os
import re
import subprocess

from grl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "grl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from grl_b200 import build
        build.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (grl_[a-z0-9_]+)", out))
    declared = header_functions()
    assert declared, "no declarations parsed"
    missing = [f for f in declared if f not in exported]
    assert not missing, "declared in include/grl_b200.h but not exported: %s" % missing
    # and the ctypes table binds exactly the declared surface
    assert sorted(_lib.exported_symbols()) == declared
    _lib.load_library()          # dlopen + signature attach, no compute


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "grl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f
