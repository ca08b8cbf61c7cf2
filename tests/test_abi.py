"""The C-ABI library loads without a GPU and exports every symbol include/riptrm_b200.h declares
(no compute calls here: kernels need a device)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "riptrm_b200.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    import riptrm_b200
    return riptrm_b200


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(riptrm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built._lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/riptrm_b200.h but not exported"
    # the ctypes table binds exactly the declared functions
    assert sorted(built._lib.SYMBOLS) == declared


def test_abi_version_and_struct_layout(built):
    lib = built.load_library()
    assert lib.riptrm_abi_version() == 2
    # riptrm_options: 8 x int32, 13 x double, 3 pointers; ABI 2 appends 2 x int32, 1 double, 1 pointer
    assert ctypes.sizeof(built._lib.RiptrmOptions) == 8 * 4 + 13 * 8 + 3 * 8 + 2 * 4 + 8 + 8
    src = open(HEADER).read()
    assert int(re.search(r"#define RIPTRM_TRACE_FIELDS (\d+)", src).group(1)) == built._lib.TRACE_FIELDS
    assert int(re.search(r"#define RIPTRM_SUMMARY_FIELDS (\d+)", src).group(1)) == built._lib.SUMMARY_FIELDS
    # field order of the trace / summary enums matches the Python tables
    tr = re.findall(r"RIPTRM_TR_([A-Z_]+) = (\d+)", src)
    assert len(tr) == built._lib.TRACE_FIELDS and [int(v) for _, v in tr] == list(range(len(tr)))
    sm = re.findall(r"RIPTRM_SM_([A-Z_]+) = (\d+)", src)
    assert len(sm) == built._lib.SUMMARY_FIELDS and [int(v) for _, v in sm] == list(range(len(sm)))


def test_errors_are_codes_not_exceptions(built):
    """Argument validation happens before any CUDA call: negative codes + a message, no crash."""
    lib = built.load_library()
    h = ctypes.c_void_p()
    assert lib.riptrm_create(1, 0, 1, 0, 1, 0, ctypes.byref(h)) == -1
    assert b"positive" in lib.riptrm_last_error()
    assert lib.riptrm_create(99, 5, 1, 5, 1, 0, ctypes.byref(h)) in (-1, -3)
    # shape rules of the large-n families (4 = COLUMNS, 5 = STIEFEL): batch 1, m = n p, p <= 16, Stiefel also p <= n
    assert lib.riptrm_create(5, 100, 4, 400, 2, 0, ctypes.byref(h)) == -1 and b"batch must be 1" in lib.riptrm_last_error()
    assert lib.riptrm_create(5, 100, 4, 100, 1, 0, ctypes.byref(h)) == -1 and b"m == n * p" in lib.riptrm_last_error()
    assert lib.riptrm_create(5, 100, 17, 1700, 1, 0, ctypes.byref(h)) == -3
    assert lib.riptrm_create(5, 3, 4, 12, 1, 0, ctypes.byref(h)) == -3
    assert lib.riptrm_create(4, 100, 17, 1700, 1, 0, ctypes.byref(h)) == -3
    assert lib.riptrm_destroy(None) == 0
    assert lib.riptrm_launch_count(None) == 0
    with pytest.raises(built.RiptrmError):
        built._lib.check(-1)


def test_missing_library_fails_loudly(built, monkeypatch):
    """No CPU fallback: a missing .so raises instead of routing elsewhere."""
    monkeypatch.setattr(built._lib, "_lib", None)
    monkeypatch.setattr(built._lib, "LIB_PATH", "/nonexistent/libriptrm_b200.so")
    with pytest.raises(built.RiptrmError, match="no CPU fallback"):
        built._lib.load_library()
