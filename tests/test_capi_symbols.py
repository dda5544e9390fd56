"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/voxeltoy_b200.h declares; compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from voxeltoy_b200 import build
    build.build()
    import voxeltoy_b200 as vt
    return vt.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "voxeltoy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(vt_[a-z_0-9]+)\s*\(", text)
    return sorted(set(n for n in names if n not in ("vt_log_fn",)))


def test_header_symbols_all_exported(lib):
    import voxeltoy_b200 as vt
    names = _declared()
    assert len(names) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", vt.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    assert sorted(vt.SIGNATURES) == names          # the ctypes table covers exactly the header


def test_no_cpu_fallback_without_device(lib):
    import voxeltoy_b200 as vt
    if lib.vt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    assert lib.vt_create(0, ctypes.byref(h)) == -3 and not h.value
    with pytest.raises(vt.VtError):
        vt.Context(0)


def test_product_never_touches_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "voxeltoy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                t = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|libvto|vto_", t):
                    bad.append(f)
    assert not bad, bad
