"""ctypes binding of the C ABI in include/litho_b200.h.

The product path loads ``liblitho_b200.so`` (sm_100a device code, built in-tree by
``__graft_entry__.build()`` / ``csrc/Makefile``) and fails loudly when it is missing:
there is no CPU fallback.  ``NativeLib(path)`` with an explicit path exists so that the
tests can point the very same binding at the CPU emulation of the kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
DEVICE_LIB = os.path.join(_HERE, "liblitho_b200.so")


class LithoError(RuntimeError):
    pass


class PlanInfo(C.Structure):
    _fields_ = [("pn", C.c_int), ("N", C.c_int), ("bbox", C.c_int * 4), ("L", C.c_int), ("M", C.c_int),
                ("R", C.c_int), ("Wr", C.c_int), ("path", C.c_int), ("default_batch", C.c_int),
                ("intensity_elems", C.c_uint64), ("shift_range", C.c_int * 4)]

PLAN_GENERIC = 1
RIM_LINES = 3  # LITHO_RIM_LINES
PHASE_INPUTS_READY = 4  # litho_abbe_fft_accumulate_ex: inputs valid on the device at call time


# every symbol include/litho_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("litho_abi_version", C.c_int, []),
    ("litho_last_error", C.c_char_p, []),
    ("litho_is_device_build", C.c_int, []),
    ("litho_epsilon_n", C.c_int, [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    ("litho_pupil_bbox", C.c_int, [_P, C.c_int, C.POINTER(C.c_int), _P]),
    ("litho_pupil_support", C.c_int, [_P, C.c_int, C.POINTER(C.c_int), _P]),
    ("litho_pupil_support_lines", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), _P]),
    ("litho_shift_bounds", C.c_int, [_P, C.c_int, C.POINTER(C.c_int), _P]),
    ("litho_source_points", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_int), _P]),
    ("litho_plan_create", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    ("litho_plan_create_ex", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    ("litho_plan_create_lines", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(_P)]),
    ("litho_plan_finalize_workspace_bytes", C.c_size_t, [_P]),
    ("litho_plan_destroy", None, [_P]),
    ("litho_plan_get_info", C.c_int, [_P, C.POINTER(PlanInfo)]),
    ("litho_plan_workspace_bytes", C.c_size_t, [_P, C.c_int]),
    ("litho_plan_column_tile", C.c_int, [_P]),
    ("litho_plan_status", C.c_int, [_P, C.POINTER(C.c_int), _P]),
    ("litho_abbe_fft_accumulate", C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    ("litho_abbe_fft_accumulate_ex", C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, C.c_size_t, _P, C.c_int]),
    ("litho_plan_workspace_bytes_focus", C.c_size_t, [_P, C.c_int, C.c_int]),
    ("litho_abbe_fft_accumulate_focus", C.c_int, [_P, _P, _P, C.c_int, C.c_size_t, _P, _P, C.c_int, C.c_int, _P, C.c_size_t,
                                                  _P, C.c_size_t, _P]),
    ("litho_mask_spectrum_workspace_bytes", C.c_size_t, [C.c_int, C.c_double, C.c_int]),
    ("litho_mask_spectrum", C.c_int, [_P, C.c_int, C.c_double, C.c_int, _P, _P, C.c_size_t, _P]),
    ("litho_direct_operator", C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int, _P, _P]),
    ("litho_direct_default_batch", C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    ("litho_direct_workspace_bytes", C.c_size_t, [C.c_int, C.POINTER(C.c_int), C.c_int]),
    ("litho_direct_accumulate", C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_int), _P, _P, C.c_int, C.c_int, _P, _P,
                                          C.c_size_t, _P]),
    ("litho_direct_status", C.c_int, [C.POINTER(C.c_int), _P]),
    ("litho_direct_field", C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_int), _P, _P, C.c_size_t, _P]),
    ("litho_direct_mask_spectrum", C.c_int, [_P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    ("litho_source_build", C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, _P, _P]),
    ("litho_pupil_build", C.c_int, [C.POINTER(C.c_float), C.c_int, C.c_int, _P, _P, _P]),
    ("litho_fp32_probe", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double), _P]),
    ("litho_fft_output_side", C.c_int, [C.c_int, C.c_double]),
    ("litho_abbe_fft_finalize", C.c_int, [_P, _P, C.c_double, _P, _P, C.c_size_t, _P]),
    ("litho_abbe_fft_unpermute", C.c_int, [_P, _P, _P, _P, C.c_size_t, _P]),
    ("litho_fft_field", C.c_int, [_P, _P, _P, _P, _P, C.c_size_t, _P]),
    ("litho_peer_alloc", C.c_int, [C.c_size_t, C.POINTER(_P), C.c_char_p]),
    ("litho_peer_open", C.c_int, [C.c_char_p, C.POINTER(_P)]),
    ("litho_peer_close", C.c_int, [_P]),
    ("litho_peer_free", C.c_int, [_P]),
    ("litho_peer_signal", C.c_int, [C.POINTER(_P), C.c_int, C.c_uint64, _P]),
    ("litho_peer_wait", C.c_int, [_P, C.c_int, C.c_uint64, _P, _P]),
    ("litho_peer_copy", C.c_int, [_P, _P, C.c_size_t, _P]),
    ("litho_peer_sum", C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_uint64, _P, C.c_uint64, _P, _P]),
    ("litho_peer_last_error", C.c_char_p, []),
]
MAX_PEERS = 16           # LITHO_MAX_PEERS
PEER_HANDLE_BYTES = 64   # LITHO_PEER_HANDLE_BYTES


class NativeLib:
    """Loaded shared library with typed entry points."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise LithoError(
                f"native library not found: {path}. Build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` or `make -C lithographysimulator_b200/csrc`. There is no CPU fallback.")
        self.path = path
        self.lib = C.CDLL(path)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(self.lib, name)  # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)
        if self.litho_abi_version() != 1:
            raise LithoError("ABI version mismatch")

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            msg = self.litho_last_error()
            raise LithoError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

    def check_peer(self, rc: int, what: str = ""):
        if rc != 0:
            msg = self.litho_peer_last_error()
            raise LithoError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

    # ---- peer memory (multi-GPU sum of the partial planes) ------------------------------------
    def peer_alloc(self, nbytes: int):
        """(pointer, handle bytes) of a zeroed buffer that the other processes of this box can map."""
        ptr = _P()
        handle = C.create_string_buffer(PEER_HANDLE_BYTES)
        self.check_peer(self.litho_peer_alloc(nbytes, C.byref(ptr), handle), "litho_peer_alloc")
        return int(ptr.value), bytes(handle.raw)

    def peer_open(self, handle: bytes) -> int:
        ptr = _P()
        self.check_peer(self.litho_peer_open(C.create_string_buffer(handle, PEER_HANDLE_BYTES), C.byref(ptr)), "litho_peer_open")
        return int(ptr.value)

    def peer_signal(self, flag_ptrs, value: int, stream: int = 0):
        arr = (_P * len(flag_ptrs))(*flag_ptrs)
        self.check_peer(self.litho_peer_signal(arr, len(flag_ptrs), value, stream), "litho_peer_signal")

    def peer_wait(self, flags_ptr: int, n: int, value: int, err_ptr=None, stream: int = 0):
        self.check_peer(self.litho_peer_wait(flags_ptr, n, value, err_ptr, stream), "litho_peer_wait")

    def peer_copy(self, dst_ptr: int, src_ptr: int, nbytes: int, stream: int = 0):
        self.check_peer(self.litho_peer_copy(dst_ptr, src_ptr, nbytes, stream), "litho_peer_copy")

    def peer_sum(self, out_ptr: int, plane_ptrs, elems: int, flags_ptr=None, value: int = 0, err_ptr=None, stream: int = 0):
        arr = (_P * len(plane_ptrs))(*plane_ptrs)
        self.check_peer(self.litho_peer_sum(out_ptr, arr, len(plane_ptrs), elems, flags_ptr, value, err_ptr, stream),
                        "litho_peer_sum")

    # ---- thin typed helpers (pointers are plain ints) --------------------------------------
    def epsilon_n(self, deltaK: float, pixelSize: float, wavelength: float):
        eps = C.c_double()
        n = C.c_int()
        self.check(self.litho_epsilon_n(deltaK, pixelSize, wavelength, C.byref(eps), C.byref(n)), "litho_epsilon_n")
        return eps.value, n.value

    def pupil_bbox(self, pupil_ptr: int, pn: int, stream: int = 0):
        box = (C.c_int * 4)()
        self.check(self.litho_pupil_bbox(pupil_ptr, pn, box, stream), "litho_pupil_bbox")
        return tuple(box)

    def pupil_support(self, pupil_ptr: int, pn: int, stream: int = 0, lines: int = RIM_LINES):
        """bbox + extents of the `lines` outermost rows/columns of the pupil support (4 + 8*lines ints)."""
        sup = (C.c_int * (4 + 8 * lines))()
        self.check(self.litho_pupil_support_lines(pupil_ptr, pn, lines, sup, stream), "litho_pupil_support_lines")
        return tuple(sup)

    def shift_bounds(self, shifts_ptr, n_src: int, stream: int = 0):
        b = (C.c_int * 4)()
        self.check(self.litho_shift_bounds(shifts_ptr, n_src, b, stream), "litho_shift_bounds")
        return tuple(b)

    def source_points(self, ls_ptr: int, elem_size: int, is_float: bool, pn: int, rank: int, world: int, shifts_ptr,
                      capacity: int, stream: int = 0):
        """(n_all, n_mine, (d0 min, d0 max, d1 min, d1 max)) -- see litho_source_points."""
        meta = (C.c_int * 6)()
        self.check(self.litho_source_points(ls_ptr, elem_size, 1 if is_float else 0, pn, rank, world, shifts_ptr, capacity,
                                            meta, stream), "litho_source_points")
        return int(meta[0]), int(meta[1]), (int(meta[2]), int(meta[3]), int(meta[4]), int(meta[5]))

    def plan_create(self, pn: int, N: int, support, flags: int = 0) -> "Plan":
        """`support` is the 4-int bbox or the (4 + 8*lines)-int result of pupil_support()."""
        handle = _P()
        n = len(support)
        if n < 4 or (n - 4) % 8:
            raise LithoError(f"plan_create: support must hold 4 + 8*lines ints, got {n}")
        arr = (C.c_int * n)(*support)
        self.check(self.litho_plan_create_lines(pn, N, arr, (n - 4) // 8, flags, C.byref(handle)),
                   "litho_plan_create_lines")
        return Plan(self, handle)


class Plan:
    def __init__(self, lib: NativeLib, handle):
        self.lib = lib
        self.handle = handle
        info = PlanInfo()
        lib.check(lib.litho_plan_get_info(handle, C.byref(info)), "litho_plan_get_info")
        self.info = info
        self.pn, self.N, self.M, self.R, self.Wr = info.pn, info.N, info.M, info.R, info.Wr
        self.bbox = tuple(info.bbox)
        self.intensity_elems = int(info.intensity_elems)
        self.default_batch = info.default_batch
        self.path = info.path
        self.shift_range = tuple(info.shift_range)

    def workspace_bytes(self, batch: int = 0) -> int:
        return int(self.lib.litho_plan_workspace_bytes(self.handle, batch))

    def accumulate(self, maskFT, pupil, shifts, weights, n_src, batch, intensity, workspace, workspace_bytes, stream=0,
                   phases=3):
        self.lib.check(self.lib.litho_abbe_fft_accumulate_ex(self.handle, maskFT, pupil, shifts, weights, n_src, batch,
                                                             intensity, workspace, workspace_bytes, stream, phases),
                       "litho_abbe_fft_accumulate")

    def workspace_bytes_focus(self, batch: int, n_focus: int) -> int:
        return int(self.lib.litho_plan_workspace_bytes_focus(self.handle, batch, n_focus))

    def accumulate_focus(self, maskFT, pupils, n_focus, pupil_stride, shifts, weights, n_src, batch, intensities,
                         intensity_stride, workspace, workspace_bytes, stream=0):
        self.lib.check(self.lib.litho_abbe_fft_accumulate_focus(self.handle, maskFT, pupils, n_focus, pupil_stride, shifts,
                                                                weights, n_src, batch, intensities, intensity_stride,
                                                                workspace, workspace_bytes, stream),
                       "litho_abbe_fft_accumulate_focus")

    def status(self, stream=0):
        """(shift_clamped, tile_copy_lost) sticky error words of a fast plan, read and cleared (synchronises)."""
        st = (C.c_int * 2)()
        self.lib.check(self.lib.litho_plan_status(self.handle, st, stream), "litho_plan_status")
        return int(st[0]), int(st[1])

    def column_tile(self) -> int:
        """Columns per TMA-staged tile of the column pass (0: plain-load column kernel)."""
        return int(self.lib.litho_plan_column_tile(self.handle))

    def output_side(self, eps: float) -> int:
        return int(self.lib.litho_fft_output_side(self.pn, eps))

    def shifts_fit(self, bounds) -> bool:
        """True if shift bounds {d0 min,max,d1 min,max} keep roll() from wrapping the pupil window."""
        r = self.shift_range
        return bounds[0] >= r[0] and bounds[1] <= r[1] and bounds[2] >= r[2] and bounds[3] <= r[3]

    def finalize_workspace_bytes(self) -> int:
        return int(self.lib.litho_plan_finalize_workspace_bytes(self.handle))

    def finalize(self, intensity, eps, out, workspace=None, workspace_bytes=0, stream=0):
        self.lib.check(self.lib.litho_abbe_fft_finalize(self.handle, intensity, eps, out, workspace, workspace_bytes,
                                                        stream), "litho_abbe_fft_finalize")

    def unpermute(self, intensity, out, workspace=None, workspace_bytes=0, stream=0):
        self.lib.check(self.lib.litho_abbe_fft_unpermute(self.handle, intensity, out, workspace, workspace_bytes,
                                                         stream), "litho_abbe_fft_unpermute")

    def fft_field(self, pf, maskFT, field, workspace, workspace_bytes, stream=0):
        self.lib.check(self.lib.litho_fft_field(self.handle, pf, maskFT, field, workspace, workspace_bytes, stream),
                       "litho_fft_field")

    def close(self):
        if self.handle:
            self.lib.litho_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_lock = threading.Lock()
_device_lib: NativeLib | None = None


def device_lib() -> NativeLib:
    """The sm_100a library.  Raises LithoError if it has not been built (no fallback)."""
    global _device_lib
    with _lock:
        if _device_lib is None:
            lib = NativeLib(DEVICE_LIB)
            if not lib.litho_is_device_build():
                raise LithoError(f"{DEVICE_LIB} is not a device build")
            _device_lib = lib
        return _device_lib
