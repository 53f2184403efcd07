"""The C-ABI library loads and exports every symbol include/zksc.h declares (no compute calls)."""
import ctypes
import os
import re

import zk_cryptography_b200 as zk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zksc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zksc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built):
    L = ctypes.CDLL(zk._lib.lib_path())
    names = declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_covers_the_header(built):
    bound = set(zk.lib()._zksc_signatures)
    assert set(declared_symbols()) == bound


def test_version_and_device_count(built):
    L = zk.lib()
    assert b"sm_100a" in L.zksc_version()
    assert L.zksc_device_count() >= 0


def test_rust_ffi_is_in_sync_with_the_header(built):
    """rust/zksc-sys/src/ffi.rs is generated from include/zksc.h (tools/gen_rust_ffi.py): every declared entry point must
    be declared there too (the Rust crates are source only -- no toolchain in the image -- so drift would go unnoticed)."""
    ffi = open(os.path.join(ROOT, "rust", "zksc-sys", "src", "ffi.rs")).read()
    rust = set(re.findall(r"pub fn (zksc_[a-z0-9_]+)\(", ffi))
    assert rust == set(declared_symbols())


def test_headers_compile_cleanly():
    """include/zksc.h is plain C (C99, no CUDA or C++ types in the signatures); include/zksc.hpp and the restated reference tests
    compile without warnings under -Wall -Wextra -Wpedantic."""
    import shutil
    import subprocess
    import pytest
    if not shutil.which("gcc") or not shutil.which("g++"):
        pytest.skip("no host compiler")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "zksc.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-fsyntax-only", os.path.join(ROOT, "tests", "cpp", "reference_cases.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
