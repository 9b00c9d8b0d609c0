"""The C-ABI library loads without a GPU, exports every symbol include/blz_cull.h declares, and FAILS LOUDLY (no CPU path)
when there is no CUDA device.  CPU only: no compute call is made."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "blz_cull.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(blz_cull_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_are_exported(built):
    from blitzen_b200 import capi
    lib = C.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/blz_cull.h but not exported: {missing}"


def test_abi_version_and_struct_sizes(built):
    from blitzen_b200 import capi
    lib = capi.load_library()
    assert lib.blz_cull_abi_version() == 1
    assert C.sizeof(capi.SceneDesc) == 168 and C.sizeof(capi.Outputs) == 160     # sizeof(blz_scene_desc) / sizeof(blz_outputs) on LP64


def test_no_device_is_an_error_not_a_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from blitzen_b200 import capi
    with pytest.raises(capi.BlzError) as e:
        capi.CullContext(0)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "blitzen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "cull_oracle" not in text, f"{f} references the oracle"
