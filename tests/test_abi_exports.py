"""The C-ABI library loads and exports every symbol include/twxi.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "twxi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(twxi_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from topowx_b200 import build
    lib_path = build.build_lib()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libtwxi.so does not export %s" % n


def test_binding_covers_header():
    from topowx_b200 import _lib
    assert set(_declared()) == set(_lib.EXPORTED)
    assert _lib.lib.twxi_version() == 100


def test_no_gpu_is_an_error_not_a_fallback():
    """Without a CUDA device the product must fail loudly (there is no CPU path behind the API)."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from topowx_b200 import synth, db, _lib
    from topowx_b200.context import TwxiContext
    da = synth.make_station_db(0, 150, synth.tile_bbox(buf=1.0), synth.Fields(), synth.make_days(1995, 1))
    with pytest.raises(_lib.TwxiError):
        TwxiContext(da, np.isnan(da.stns[db.BAD]))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "topowx_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_header_is_plain_c():
    """include/twxi.h is the C ABI: it must compile as C (no C++ constructs, no CUDA / torch types)."""
    import shutil
    import subprocess
    import pytest
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = '#include "twxi.h"\nint main(void) { twxi_points p; p.npts = 0; return (int)sizeof(p) * 0 + TWXI_OK; }\n'
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-"], input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr
