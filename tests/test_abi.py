"""CPU: the C-ABI library loads and exports every symbol include/b200vc.h declares, with the argument counts
the ctypes binding assumes.  No compute entry point is exercised here (no GPU in the CPU suite) except to
check that argument validation / error reporting works without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, "include", "b200vc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"B200VC_API\s+([\w\s\*]+?)\s*\b(b200vc_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(2)] = n
    return out


def test_header_declares_the_bound_set():
    from b200vc import _lib
    decl = declared()
    assert set(decl) == set(_lib.EXPORTS), set(decl) ^ set(_lib.EXPORTS)
    for name, n in decl.items():
        assert len(_lib._SIGNATURES[name][1]) == n, (name, n, len(_lib._SIGNATURES[name][1]))


def test_library_exports_every_declared_symbol():
    from b200vc import _lib
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared():
        assert hasattr(raw, name), f"{name} is declared in include/b200vc.h but not exported"
    assert lib.b200vc_version() == 100
    assert lib.b200vc_reduce_blocks(1) == 1
    assert lib.b200vc_reduce_blocks(1044480) == 1020
    assert lib.b200vc_reduce_blocks(10 ** 9) == 1184
    assert lib.b200vc_gdn_params_floats(128) == 128 + 4 * 128 * 128


def test_argument_validation_reports_errors_without_a_device():
    from b200vc import _lib
    lib = _lib.load()
    rc = lib.b200vc_warp_f32(None, 0, None, None, None, None, 0, 1, 3, 8, 8, 0, 0, None)
    assert rc == -1 and "null pointer" in _lib.last_error()
    rc = lib.b200vc_gdn_f32(None, None, None, None, 1, 128, 64, 0, 0, None)
    assert rc == -1
    with pytest.raises(RuntimeError, match="gdn_f32"):
        _lib.check(rc, "gdn_f32")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from b200vc import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or eager-PyTorch fallback"):
        _lib.load()
