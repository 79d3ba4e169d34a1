"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/wrach_cuda.h declares, the uniform is 32 bytes with the reference's offsets, and without a
GPU the worker refuses to exist (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wrach_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wrach_cuda_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from wrach_b200 import _ffi
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), "libwrach_cuda.so does not export %s" % name
    assert set(names) == set(_ffi.SYMBOLS), "ctypes table and header disagree"
    assert _ffi.lib().wrach_cuda_version().startswith(b"wrach_cuda sm_100a")


def test_uniform_layout_matches_the_reference():
    from wrach_b200 import WorldSettings
    assert ctypes.sizeof(WorldSettings) == 32  # config_shader.rs:15-29
    offs = {n: getattr(WorldSettings, n).offset for n, _ in WorldSettings._fields_}
    assert offs == {"view_dimensions": 0, "view_anchor": 8, "grid_dimensions": 16, "cell_size": 24,
                    "particles_in_frame_count": 28}


def test_strip_column_split_is_exhaustive():
    from wrach_b200 import _ffi
    L = _ffi.lib()
    for gx, n in ((21845, 8), (1822, 4), (7, 8), (456, 2), (5, 1)):
        covered, prev_end = 0, 0
        for r in range(n):
            b, e = ctypes.c_uint32(), ctypes.c_uint32()
            L.wrach_cuda_strip_columns(gx, r, n, ctypes.byref(b), ctypes.byref(e))
            assert b.value == prev_end and e.value >= b.value
            covered += e.value - b.value
            prev_end = e.value
        assert covered == gx


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import wrach_b200
    s = wrach_b200.WorldSettings()
    s.view_dimensions[:] = [10.0, 10.0]
    s.grid_dimensions[:] = [4, 4]
    s.cell_size = 3
    with pytest.raises(wrach_b200.WrachCudaError) as e:
        wrach_b200.PhysicsComputeWorker(s, 18, 164)
    assert e.value.status == -3 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wrach_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.lower(), "%s mentions the oracle" % os.path.join(dirpath, fn)


def test_headers_are_plain_c_and_the_c_example_links(tmp_path):
    """include/*.h compile as strict C11 and examples/api_smoke.c -- the reference's API smoke test
    (runners/api/src/lib.rs:102-126) written against nothing but the two headers -- links against the
    library.  Without a device it must stop at wrach_api_new with the no-fallback message (exit 3);
    with one it must pass (exit 0)."""
    import subprocess
    import torch
    exe = str(tmp_path / "api_smoke")
    libdir = os.path.join(ROOT, "wrach_b200", "lib")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "api_smoke.c"), "-L" + libdir, "-lwrach_cuda",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
        assert "read back 164 positions" in r.stdout
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


def test_rust_binding_declares_every_entry_point():
    """ffi/rust/wrach-cuda cannot be compiled here (no Rust toolchain); at least its extern block must
    name exactly the entry points the header declares, with the uniform at 32 bytes."""
    src = open(os.path.join(ROOT, "ffi", "rust", "wrach-cuda", "src", "lib.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n    }\n")]
    rust = sorted(set(re.findall(r"pub fn (wrach_cuda_[a-z_0-9]+)\s*\(", block)))
    assert rust == declared_symbols()
    assert "size_of::<WorldSettings>() == 32" in src
