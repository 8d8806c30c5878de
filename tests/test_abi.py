"""CPU: the C-ABI library builds, loads, exports every symbol include/b2s.h declares, and fails loudly without a GPU."""
import os
import re

import pytest

from calibrating_b200 import _ffi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _ffi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2s.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libb2s.so does not export %s" % n
    assert sorted(_ffi.SIGNATURES) == names, "ctypes stub and header disagree"


def test_struct_layouts_match_header():
    import ctypes
    assert ctypes.sizeof(_ffi.SgbmParams) == 12 * 4
    assert ctypes.sizeof(_ffi.DepthOut) == 8 * 8
    assert ctypes.sizeof(_ffi.Timing) == 7 * 4 + 2 * 4
    assert ctypes.sizeof(_ffi.Rig) == 6 * 4 + 9 * 8 + 5 * 8 + 2 * 4


def test_struct_sizes_against_the_c_compiler(tmp_path):
    """sizeof of every struct of include/b2s.h as gcc sees it == the ctypes mirror."""
    import ctypes
    import subprocess
    names = {"b2s_sgbm_params": _ffi.SgbmParams, "b2s_rig": _ffi.Rig, "b2s_map_params": _ffi.MapParams, "b2s_rig_params": _ffi.RigParams,
             "b2s_depth_out": _ffi.DepthOut, "b2s_timing": _ffi.Timing}
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "b2s.h"\nint main(void){' +
                   "".join('printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, t in names.items():
        assert int(out[n]) == ctypes.sizeof(t), n


def test_no_cpu_fallback(lib):
    if lib.b2s_device_count() > 0:
        pytest.skip("a GPU is present")
    import calibrating_b200 as cb
    with pytest.raises(_ffi.B2SError, match="no CPU fallback"):
        cb.StereoSGBM_create(numDisparities=16)
    with pytest.raises(_ffi.B2SError):
        cb.SemiGlobalBlockMatching()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "calibrating_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
