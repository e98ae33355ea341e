"""The C ABI boundary without a GPU: every function include/litho_b200.h declares is bound by the ctypes layer and
exported by both builds of the library (the sm_100a product and the CPU emulation used by the tests); host-only
entry points give the reference's numbers.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import helpers as H
from lithographysimulator_b200 import _native

HEADER = os.path.join(H.ROOT, "include", "litho_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(litho_[a-z0-9_]+)\s*\(", src)))


def _device_lib_path():
    if not os.path.exists(_native.DEVICE_LIB):   # nvcc cross-compiles without a GPU
        subprocess.run(["make", "-C", os.path.join(H.ROOT, "lithographysimulator_b200", "csrc"), "all", "-j8"],
                       check=True, stdout=subprocess.DEVNULL)
    return _native.DEVICE_LIB


def test_binding_covers_every_declared_function():
    declared = _declared()
    bound = sorted(name for name, _, _ in _native.SYMBOLS)
    assert len(declared) >= 30
    assert declared == bound, (set(declared) ^ set(bound))


@pytest.mark.parametrize("which", ["device", "emu"])
def test_library_exports_every_declared_function(which):
    path = _device_lib_path() if which == "device" else H.emu_lib().path
    lib = C.CDLL(path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.litho_is_device_build.restype = C.c_int
    assert lib.litho_is_device_build() == (1 if which == "device" else 0)


def test_host_only_entry_points_of_the_device_build():
    lib = _native.NativeLib(_device_lib_path())
    assert lib.litho_abi_version() == 1
    eps, N = lib.epsilon_n(4 / 64, 25.0, 193.0)            # (1.0362694300518134, 128)
    assert N == 128 and abs(eps - 1.0362694300518134) < 1e-15
    assert lib.litho_fft_output_side(4096, eps) == 4094    # round/floor mismatch at 4096
    assert lib.litho_fft_output_side(2048, eps) == 2048
    # argument errors are reported through the return code + litho_last_error, never by crashing
    handle = C.c_void_p()
    bbox = (C.c_int * 4)(0, 10, 0, 10)
    assert lib.litho_plan_create_lines(63, 128, bbox, 0, 0, C.byref(handle)) != 0       # odd grid
    assert b"even" in lib.litho_last_error()
    assert lib.litho_plan_create_lines(64, 100, bbox, 0, 0, C.byref(handle)) != 0       # N not a power of two
    assert lib.litho_plan_create_lines(64, 128, bbox, 7, 0, C.byref(handle)) != 0       # lines out of range
    assert lib.litho_plan_create_lines(256, 128, bbox, 0, 0, C.byref(handle)) != 0      # N < pn (reference raises)
    assert lib.litho_plan_column_tile(None) == 0
    assert lib.litho_plan_workspace_bytes(None, 1) == 0
