"""CPU: the C-ABI library builds, loads and exports every symbol include/pvw_b200.h declares; the host layer fails
loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pvw_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pvw_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("pvw_build", os.path.join(ROOT, "pvw-rs_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "pvw_b200.h"\nint main(void){ pvw_params_desc d; (void)d; return PVW_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o", str(tmp_path / "t.o")])


def test_library_exports_every_declared_symbol(lib_path):
    names = declared_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pvw_b200.h but not exported"
    import pvw_rs_b200
    assert set(pvw_rs_b200._ffi.SIGNATURES) == set(names)
    assert b"sm_100a" in pvw_rs_b200._ffi.load().pvw_version()


def test_library_is_sm100a_sass(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import pvw_rs_b200
    with pytest.raises(pvw_rs_b200.PvwError) as ei:
        pvw_rs_b200.Engine(3, 4, 8, [0xFFFFEE001, 0xFFFFC4001])
    assert ei.value.variant == "InternalError" and "no CPU fallback" in str(ei.value)
    # parameter validation happens before the device is touched and maps to InvalidParameters (parameters.rs:131-148)
    with pytest.raises(pvw_rs_b200.PvwError) as ei:
        pvw_rs_b200.Engine(3, 4, 12, [0xFFFFEE001])
    assert ei.value.variant == "InvalidParameters"
    null = ctypes.c_void_p()
    assert pvw_rs_b200._ffi.load().pvw_ctx_synchronize(null) != 0


def test_cpp_host_mirror_compiles(lib_path):
    exe = os.path.join(ROOT, "pvw-rs_b200", "build", "pvw_example")
    assert os.path.exists(exe)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stdout      # fails loudly without a device
