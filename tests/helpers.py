"""Shared test helpers: golden loading, the CPU-emulation library, and a driver that runs the
C ABI hot path on either backend (numpy host pointers for the emulation, torch CUDA tensors for
the device build)."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from lithographysimulator_b200 import _native  # noqa: E402
from oracle import abbe_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "liblitho_emu.so")
TOL = 1e-5  # north-star tolerance: relative L2 on the aerial image


def load_kat():
    z = np.load(os.path.join(GOLDEN, "kat_small.npz"))
    cases = {}
    for key in z.files:
        case, name = key.split("/")
        cases.setdefault(case, {})[name] = z[key]
    return cases


_emu = None


def emu_lib() -> _native.NativeLib:
    """Build (once) and load the CPU emulation of the kernels.  Test infrastructure only."""
    global _emu
    if _emu is None:
        subprocess.run(["make", "-C", os.path.join(ROOT, "lithographysimulator_b200", "csrc"), "emu", "-j8"],
                       check=True, stdout=subprocess.DEVNULL)
        _emu = _native.NativeLib(EMU_LIB)
        assert _emu.litho_is_device_build() == 0
    return _emu


def _ptr(a):
    return a.ctypes.data


def emu_abbe_fft(maskFT, pupil, lightsource, pixel_size, wavelength, batch=0, weights=None, postprocess=True,
                 shifts=None, force_generic=False):
    """abbeImage(fft=True) through the C ABI on the CPU emulation (numpy buffers)."""
    lib = emu_lib()
    pn = maskFT.shape[0]
    eps, N = lib.epsilon_n(4 / pn, pixel_size, wavelength)
    maskFT = np.ascontiguousarray(maskFT, dtype=np.complex64)
    pupil = np.ascontiguousarray(pupil, dtype=np.complex64)
    support = lib.pupil_support(_ptr(pupil), pn)
    if shifts is None:
        shifts = O.source_shifts(lightsource, pn)
    shifts = np.ascontiguousarray(shifts, dtype=np.int32)
    n_src = shifts.shape[0]
    plan = lib.plan_create(pn, N, support)
    bounds = lib.shift_bounds(_ptr(shifts) if n_src else None, n_src)
    if force_generic or (plan.path == 2 and not plan.shifts_fit(bounds)):
        plan.close()
        plan = lib.plan_create(pn, N, support, _native.PLAN_GENERIC)
    inten = np.zeros(plan.intensity_elems, dtype=np.float32)
    wsb = plan.workspace_bytes(batch)
    ws = np.zeros(max(wsb, 8), dtype=np.uint8)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
    plan.accumulate(_ptr(maskFT), _ptr(pupil), _ptr(shifts) if n_src else None, None if w is None else _ptr(w),
                    n_src, batch, _ptr(inten), _ptr(ws), wsb)
    fwb = plan.finalize_workspace_bytes()
    fws = np.zeros(max(fwb, 8), dtype=np.uint8)
    if postprocess:
        side = plan.output_side(eps)
        out = np.zeros((side, side), dtype=np.float32)
        plan.finalize(_ptr(inten), eps, _ptr(out), _ptr(fws), fwb)
    else:
        out = np.zeros((pn, pn), dtype=np.float32)
        plan.unpermute(_ptr(inten), _ptr(out), _ptr(fws), fwb)
    info = dict(N=N, eps=eps, M=plan.M, R=plan.R, path=plan.path, support=support, n_src=n_src, bounds=bounds,
                column_tile=plan.column_tile())
    plan.close()
    return out, info
