"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200match.h declares,
fails loudly without a GPU (no CPU fallback), and the product never touches the oracle."""
import ctypes as C
import os
import re

import pytest

from alicevision_b200 import matching

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(header="b200match.h", prefix="b200m_"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = matching.load_library()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200match.h but not exported"
    assert sorted(matching.ABI_SYMBOLS) == names
    from alicevision_b200 import regions_io
    io_names = header_functions("b200io.h", "b200io_")
    assert len(io_names) >= 12 and sorted(regions_io.IO_SYMBOLS) == io_names
    for n in io_names:
        assert hasattr(lib, n), f"{n} declared in include/b200io.h but not exported"
    from alicevision_b200 import voctree
    voc_names = header_functions("b200voc.h", "b200v_")
    assert len(voc_names) >= 19 and sorted(voctree.VOC_SYMBOLS) == voc_names
    for n in voc_names:
        assert hasattr(lib, n), f"{n} declared in include/b200voc.h but not exported"


def test_no_gpu_fails_loudly():
    lib = matching.load_library()
    if lib.b200m_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.b200m_ctx_create(0, None, C.byref(h))
    assert rc == 2 and not h and b"no CPU path" in lib.b200m_last_error()
    with pytest.raises(matching.B200MatchError):
        matching.Context(0)


def test_product_does_not_use_oracle():
    """Nothing under alicevision_b200/ or include/ may reference the oracle (parity would be void)."""
    bad = []
    for base in ("alicevision_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"import\s+oracle|from\s+oracle|oracle/|libport_oracle|libref_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_bad_arguments_return_status():
    lib = matching.load_library()
    assert lib.b200m_match_pairs(None, None, 0, C.c_float(0.8), 0, 2, None) == 1
    assert lib.b200m_upload_view(None, 0, None, 0, 128, 1, None) == 1
    assert lib.b200m_result_num_pairs(None) == 0
    assert lib.b200m_version() >= 100
