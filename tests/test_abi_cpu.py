"""CPU: the C-ABI shared library loads, exports every symbol include/openstereo_b200.h declares, and the
host-side argument checks behave like the reference's (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__
    __graft_entry__.build()
    from openstereo_b200 import _lib
    return _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "openstereo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(native):
    names = declared_symbols()
    assert len(names) >= 13
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), "%s declared in the header but not exported" % name
    # every compute entry point bound by the Python layer is declared in the header
    for name in native.SIGNATURES:
        assert name in names


def test_binding_signatures_match_header(native):
    """Arity and C types of every ctypes binding are checked against the prototypes in the header."""
    text = open(os.path.join(ROOT, "include", "openstereo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    kinds = {"const float*": ctypes.c_void_p, "float*": ctypes.c_void_p, "const void*": ctypes.c_void_p, "int": ctypes.c_int, "float": ctypes.c_float,
             "osb_stream_t": ctypes.c_void_p, "long long": ctypes.c_longlong}
    for name, argtypes in native.SIGNATURES.items():
        m = re.search(r"int\s+%s\s*\(([^)]*)\)" % name, text)
        assert m, name
        params = [p.strip() for p in m.group(1).replace("\n", " ").split(",")]
        want = []
        for p in params:
            ctype = p.rsplit(" ", 1)[0].strip()
            assert ctype in kinds, (name, p)
            want.append(kinds[ctype])
        assert want == list(argtypes), "%s: header %d params vs binding %d" % (name, len(want), len(argtypes))


def test_version_and_error_channel(native):
    assert native.lib.osb_abi_version() == 1
    assert native.launch_count() >= 0
    # argument validation happens before any CUDA call: a null pointer is OSB_EINVAL -> ValueError
    with pytest.raises(ValueError, match="null pointer"):
        native.call("osb_gwc_volume_fwd", None, None, None, 1, 8, 4, 4, 4, 2, None)


def test_ops_refuse_cpu_tensors(native):
    from openstereo_b200 import ops
    x = torch.randn(1, 8, 4, 16)
    with pytest.raises(RuntimeError, match="not implemented on the CPU"):
        ops.build_gwc_volume(x, x, 4, 2)
    with pytest.raises(RuntimeError, match="not implemented on the CPU"):
        ops.softargmin(torch.randn(1, 4, 4, 4), 4)
    with pytest.raises(ValueError, match="expected 4D input"):
        ops.faster_soft_argmin(torch.randn(4, 4, 4), 4)          # psmnet_disp_processor.py:56-58
    with pytest.raises(NotImplementedError):
        ops.cat_fms(x, x, max_disp=4, start_disp=-1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openstereo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
