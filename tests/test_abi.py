"""CPU-only: the C-ABI library loads and exports every symbol include/tecdsa_b200.h declares;
argument validation that needs no GPU."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(entry.LIB):
        entry.build()
    return ctypes.CDLL(entry.LIB)


def _declared():
    src = open(os.path.join(entry.ROOT, "include", "tecdsa_b200.h")).read()
    return sorted(set(re.findall(r"\b(tecdsa_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(lib, pkg):
    names = _declared()
    assert names, "no declarations found"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(pkg.EXPORTS) == names


def test_product_never_imports_oracle():
    """The product path must not route through the oracle or any CPU fallback."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle)|libgg20_ref|gg20_ref\.c", re.M)
    for base, _, files in os.walk(entry.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert not pat.search(open(os.path.join(base, f)).read()), f


def test_null_ctx_is_an_error_not_a_crash(lib):
    lib.tecdsa_last_error.restype = ctypes.c_char_p
    assert lib.tecdsa_ctx_sync(None) < 0
    assert lib.tecdsa_modexp_batch(None, 2048, 64, None, None, None, None, ctypes.c_size_t(0), None, None, ctypes.c_size_t(1), 0) < 0
    assert b"null" in lib.tecdsa_last_error()


def test_no_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.EngineError):
        pkg.Engine(0)


def test_limb_packing_roundtrip(pkg):
    vals = [0, 1, 2**32 - 1, 2**2047 + 12345, 2**2048 - 1]
    assert pkg.limbs_to_ints(pkg.ints_to_limbs(vals, 64)) == vals


def _build_c_consumer(tmp_path):
    """tests/c/abi_smoke.c: a C99 program over the header and the shared library, no Python in the process"""
    import subprocess
    exe = os.path.join(str(tmp_path), "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(entry.ROOT, "include"),
                           os.path.join(entry.ROOT, "tests", "c", "abi_smoke.c"), "-L", entry.PKG_DIR, "-ltecdsa_b200",
                           "-Wl,-rpath," + entry.PKG_DIR, "-o", exe])
    return exe


def test_header_is_plain_c_and_library_links_without_python(lib, tmp_path):
    import subprocess
    import torch
    exe = _build_c_consumer(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, r.stdout + r.stderr          # loaded, exported everything it needed, refused to run without a GPU


@pytest.mark.gpu
def test_c_consumer_on_gpu(lib, tmp_path):
    import subprocess
    r = subprocess.run([_build_c_consumer(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "abi_smoke ok" in r.stdout, r.stdout + r.stderr


def test_rust_ffi_lists_every_export():
    """bindings/rust/src/ffi.rs is generated from the header (tools/gen_rust_ffi.py): every declared entry point has its
    `extern "C"` declaration, and nothing else"""
    src = open(os.path.join(entry.ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (tecdsa_[a-z0-9_]+)\(", src)))
    assert rust == _declared()
