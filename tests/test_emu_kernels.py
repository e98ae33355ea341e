"""Runs the CUDA kernel bodies thread-for-thread on the CPU (tests/emu) through the C ABI and
checks them against the reference goldens and the oracle.  This validates the index arithmetic
of the device code without a GPU; it is not a product path (the product loads the sm_100a
library only and has no CPU fallback)."""
import numpy as np
import pytest

import helpers as H
from oracle import abbe_oracle as O

KAT = H.load_kat()


@pytest.mark.parametrize("name", ["demo64_quasar", "ps50_64", "ps12_64", "np2_96", "wrap_128", "shifted_128",
                                  "dense_64"])
def test_emu_fft_image_matches_reference(name):
    c = KAT[name]
    img, info = H.emu_abbe_fft(c["maskFT"], c["pupil"], c["lightsource"], float(c["pixel_size"]), 193.0)
    assert img.shape == c["image"].shape
    assert O.rel_l2(img, c["image"]) < H.TOL


def test_emu_batching_weights_and_empty_source():
    c = KAT["demo64_quasar"]
    shifts = O.source_shifts(c["lightsource"], 64)[:23]
    rng = np.random.default_rng(0)
    w = rng.uniform(0.5, 2.0, len(shifts)).astype(np.float32)
    a, _ = H.emu_abbe_fft(c["maskFT"], c["pupil"], None, 25.0, 193.0, batch=1, weights=w, shifts=shifts, postprocess=False)
    b, _ = H.emu_abbe_fft(c["maskFT"], c["pupil"], None, 25.0, 193.0, batch=7, weights=w, shifts=shifts, postprocess=False)
    assert O.rel_l2(a, b) < 1e-6
    ref = np.zeros((64, 64))
    for (d0, d1), wi in zip(shifts, w):
        e = O.calculate_fft_aerial(np.roll(c["pupil"], (d0, d1), (0, 1)), c["maskFT"], 64, 128)
        ref += wi * np.abs(e) ** 2
    assert O.rel_l2(a, ref) < H.TOL
    z, _ = H.emu_abbe_fft(c["maskFT"], c["pupil"], np.zeros((64, 64), np.int64), 25.0, 193.0)
    assert z.shape == (64, 64) and not z.any()


@pytest.mark.parametrize("pn,ps,win", [(128, 25, 33), (128, 12, 33), (64, 25, 30), (128, 25, 65)])
def test_emu_fast_path_dense_window(pn, ps, win):
    """Fast coarse-grid path on a dense random window.  win = M+1 makes the rim rows/columns fully
    populated, so the aliased +-M frequency line carries real energy and the rim sums must be right;
    win < M+1 exercises the no-rim branch.  Checked against the oracle and the generic path."""
    rng = np.random.default_rng(pn + win)
    lo = pn // 2 - win // 2
    pup = np.zeros((pn, pn), np.complex64)
    pup[lo:lo + win, lo:lo + win] = rng.standard_normal((win, win)) + 1j * rng.standard_normal((win, win))
    mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
    shifts = np.array([[0, 0], [5, -7], [-lo, lo - 1 + (pn - 2 * lo - win + 1)], [3, 3]], np.int32)
    shifts[2] = [-lo, pn - (lo + win)]  # extreme corner of the no-wrap range
    w = np.array([1.0, 2.0, 0.5, 1.5], np.float32)
    fast, info = H.emu_abbe_fft(mft, pup, None, float(ps), 193.0, shifts=shifts, weights=w, postprocess=False, batch=3)
    assert info["path"] == 2, info
    _, N = O.calculate_epsilon_n(4 / pn, ps, 193.0)
    ref = np.zeros((pn, pn))
    for (d0, d1), wi in zip(shifts, w):
        ref += wi * np.abs(O.calculate_fft_aerial(np.roll(pup, (d0, d1), (0, 1)), mft, pn, N)) ** 2
    assert O.rel_l2(fast, ref) < H.TOL
    gen, info2 = H.emu_abbe_fft(mft, pup, None, float(ps), 193.0, shifts=shifts, weights=w, postprocess=False,
                                force_generic=True)
    assert info2["path"] == 1
    assert O.rel_l2(fast, gen) < H.TOL


@pytest.mark.parametrize("wr,wc", [(34, 34), (35, 35), (35, 33), (31, 34), (33, 35), (1, 35), (34, 2)])
@pytest.mark.parametrize("tma", ["1", "0"])
def test_emu_fast_path_window_beyond_even_fit(monkeypatch, wr, wc, tma):
    """Windows of M+2 / M+3 samples (M = 32), as the reference's fp16 pupil grid produces at pn = 8192 (support
    pn/2+3): the inputs beyond M fold onto the first slots and the frequency lines M .. S-1, which alias on the
    coarse grid, come from the generalised rim sums (row/column pairs, aliased partners subtracted from the interior
    bins).  Dense random windows make every one of those lines carry energy; rows and columns may differ."""
    monkeypatch.setenv("LITHO_TMA", tma)
    pn = 128
    rng = np.random.default_rng(1000 + 10 * wr + wc)
    r0, c0 = pn // 2 - wr // 2, pn // 2 - wc // 2
    pup = np.zeros((pn, pn), np.complex64)
    pup[r0:r0 + wr, c0:c0 + wc] = rng.standard_normal((wr, wc)) + 1j * rng.standard_normal((wr, wc))
    mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
    shifts = np.array([[0, 0], [5, -7], [-r0, pn - (c0 + wc)], [3, 3], [-9, 4]], np.int32)
    w = np.array([1.0, 2.0, 0.5, 1.5, 0.75], np.float32)
    fast, info = H.emu_abbe_fft(mft, pup, None, 25.0, 193.0, shifts=shifts, weights=w, postprocess=False, batch=2)
    assert info["path"] == 2 and info["M"] == 32, info
    ref = np.zeros((pn, pn))
    for (d0, d1), wi in zip(shifts, w):
        ref += wi * np.abs(O.calculate_fft_aerial(np.roll(pup, (d0, d1), (0, 1)), mft, pn, 256)) ** 2
    assert O.rel_l2(fast, ref) < H.TOL
    if min(wr, wc) < 30:
        return
    # sparse rim lines (only a short stretch populated, like a disc's edge) with measured extents
    pup2 = pup.copy()
    pup2[r0, :] = 0; pup2[r0, c0 + 10:c0 + 15] = 1 + 1j
    pup2[r0 + wr - 1, :] = 0; pup2[r0 + wr - 1, c0 + 20:c0 + 22] = 2 - 1j
    pup2[:, c0] = 0; pup2[r0 + 5:r0 + 9, c0] = -1j
    fast2, info2 = H.emu_abbe_fft(mft, pup2, None, 25.0, 193.0, shifts=shifts, postprocess=True, batch=3)
    assert info2["path"] == 2
    ref2 = O.abbe_image(mft, pup2, None, 25, 4 / pn, 193.0, True, np.complex128, shifts=shifts)
    assert O.rel_l2(fast2, ref2) < H.TOL


@pytest.mark.parametrize("seed", range(14))
def test_emu_random_windows_sources_and_grids(seed):
    """Seeded sweep over the plan space: grid 64/96/128 px, pixel size 12/25/50 nm (N = 4pn, 2pn, pn), pupil
    windows of any size and position (1 px up to beyond M+3, off-centre), sources inside and outside the no-wrap
    range (the latter must fall back to the generic kernels), with and without weights and post-processing --
    each against the oracle's literal restatement of the reference loop."""
    rng = np.random.default_rng(4242 + seed)
    pn = int(rng.choice([64, 96, 128]))
    ps = float(rng.choice([12, 25, 50]))
    _, N = O.calculate_epsilon_n(4 / pn, ps, 193.0)
    if N < pn:          # 96 px at 50 nm: the reference itself raises there
        ps = 25.0
        _, N = O.calculate_epsilon_n(4 / pn, ps, 193.0)
    wr, wc = (int(rng.integers(1, pn // 2 + 8)) for _ in range(2))
    r0, c0 = int(rng.integers(0, pn - wr + 1)), int(rng.integers(0, pn - wc + 1))
    pup = np.zeros((pn, pn), np.complex64)
    pup[r0:r0 + wr, c0:c0 + wc] = rng.standard_normal((wr, wc)) + 1j * rng.standard_normal((wr, wc))
    if seed % 3 == 0:   # sparse rim lines, as a disc has
        pup[r0, c0 + wc // 3:] = 0
        pup[r0:r0 + wr // 2, c0 + wc - 1] = 0
    mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
    n_src = int(rng.integers(1, 6))
    if seed % 4 == 3:   # anywhere on the grid: some points wrap the window around
        shifts = rng.integers(-pn // 2, pn // 2, (n_src, 2)).astype(np.int32)
    else:               # inside the no-wrap range
        shifts = np.stack([rng.integers(-r0, pn - (r0 + wr) + 1, n_src),
                           rng.integers(-c0, pn - (c0 + wc) + 1, n_src)], 1).astype(np.int32)
    w = rng.uniform(0.25, 2.0, n_src).astype(np.float32) if seed % 2 else None
    post = bool(seed % 5 in (1, 2)) and w is None
    img, info = H.emu_abbe_fft(mft, pup, None, ps, 193.0, shifts=shifts, weights=w, postprocess=post,
                               batch=int(rng.integers(0, 4)))
    ref = np.zeros((pn, pn))
    for i, (d0, d1) in enumerate(shifts):
        wi = 1.0 if w is None else float(w[i])
        ref += wi * np.abs(O.calculate_fft_aerial(np.roll(pup, (int(d0), int(d1)), (0, 1)), mft, pn, N)) ** 2
    if post:
        eps, _ = O.calculate_epsilon_n(4 / pn, ps, 193.0)
        ref = O.fft_postprocess(ref, pn, eps, dtype=np.float64)
    assert img.shape == ref.shape, (img.shape, ref.shape, info)
    assert O.rel_l2(img, ref) < H.TOL, info


@pytest.mark.parametrize("tma", ["1", "0"])
def test_emu_column_pass_tma_and_plain_agree(monkeypatch, tma):
    """The TMA-staged column kernel (tile of source point sl+1 copied to shared memory while sl is transformed)
    and the plain-load one share the arithmetic: both must match the oracle, with batches that make the tile
    ring wrap (5 source points, batch 2 -> 3 launches over the 3 slots) and with a short window (Sr < M:
    masked rows) as well as the full rim case (Sr = M+1)."""
    monkeypatch.setenv("LITHO_TMA", tma)
    for pn, win in ((128, 33), (128, 27)):
        rng = np.random.default_rng(100 + win)
        lo = pn // 2 - win // 2
        pup = np.zeros((pn, pn), np.complex64)
        pup[lo:lo + win, lo:lo + win] = rng.standard_normal((win, win)) + 1j * rng.standard_normal((win, win))
        mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
        shifts = np.array([[0, 0], [5, -7], [-3, 9], [3, 3], [-11, 2]], np.int32)
        w = np.array([1.0, 2.0, 0.5, 1.5, 0.25], np.float32)
        img, info = H.emu_abbe_fft(mft, pup, None, 25.0, 193.0, shifts=shifts, weights=w, postprocess=False, batch=2)
        assert info["path"] == 2 and (info["column_tile"] > 0) == (tma == "1"), info
        ref = np.zeros((pn, pn))
        for (d0, d1), wi in zip(shifts, w):
            ref += wi * np.abs(O.calculate_fft_aerial(np.roll(pup, (d0, d1), (0, 1)), mft, pn, 256)) ** 2
        assert O.rel_l2(img, ref) < H.TOL


def test_emu_fast_path_subfft_256():
    """pn = 512 -> sub-FFT 256 = radix 32 x 8 (two passes with a partial second radix)."""
    c_pn = 512
    pf, _ = O.pupil_function([0, 0, 0.01, 0, 60, 0.01], c_pn, 0.7, 193.0)
    rng = np.random.default_rng(1)
    mft = (rng.standard_normal((c_pn, c_pn)) + 1j * rng.standard_normal((c_pn, c_pn))).astype(np.complex64)
    shifts = np.array([[10, -20]], np.int32)
    img, info = H.emu_abbe_fft(mft, pf, None, 25.0, 193.0, shifts=shifts, postprocess=False)
    assert info["path"] == 2 and info["M"] == 256, info
    e = O.calculate_fft_aerial(np.roll(pf, (10, -20), (0, 1)), mft, c_pn, 1024)
    assert O.rel_l2(img, np.abs(e) ** 2) < H.TOL


def test_emu_complex_field():
    f = KAT["field_fft_64"]
    lib = H.emu_lib()
    pf = np.ascontiguousarray(f["pf"])
    mft = np.ascontiguousarray(f["maskFT"])
    plan = lib.plan_create(64, 128, lib.pupil_bbox(pf.ctypes.data, 64))
    wsb = plan.workspace_bytes(1)
    ws = np.zeros(wsb, np.uint8)
    field = np.zeros((64, 64), np.complex64)
    plan.fft_field(pf.ctypes.data, mft.ctypes.data, field.ctypes.data, ws.ctypes.data, wsb)
    assert np.linalg.norm(field - f["field"]) / np.linalg.norm(f["field"]) < H.TOL


def test_emu_larger_subfft_sizes():
    """One source point through M = 128 / 256 (three radix passes are exercised by M >= 512 on the GPU)."""
    rng = np.random.default_rng(3)
    for pn, box in ((256, (64, 192)), (256, (10, 250))):
        pup = np.zeros((pn, pn), np.complex64)
        lo, hi = box
        n = hi - lo + 1
        pup[lo:hi + 1, lo:hi + 1] = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
        shifts = np.array([[3, -5]], np.int32)
        img, info = H.emu_abbe_fft(mft, pup, None, 25.0, 193.0, shifts=shifts, postprocess=False)
        e = O.calculate_fft_aerial(np.roll(pup, (3, -5), (0, 1)), mft, pn, 512)
        assert O.rel_l2(img, np.abs(e) ** 2) < H.TOL, info


@pytest.mark.parametrize("name", ["demo64_quasar", "ps50_64", "ps12_64", "np2_96", "wrap_128"])
def test_emu_mask_spectrum_matches_reference(name):
    """Mask._ffFraunhofer through litho_mask_spectrum (incl. N == pn, where the resampled mask is cropped)."""
    c = KAT[name]
    lib = H.emu_lib()
    geom = np.ascontiguousarray(c["geometry"], dtype=np.int16)
    pn = geom.shape[0]
    eps, N = float(c["eps"]), int(c["N"])
    nbytes = lib.litho_mask_spectrum_workspace_bytes(pn, eps, N)
    assert nbytes > 0
    ws = np.zeros(nbytes, np.uint8)
    out = np.zeros((pn, pn), np.complex64)
    lib.check(lib.litho_mask_spectrum(geom.ctypes.data, pn, eps, N, out.ctypes.data, ws.ctypes.data, nbytes, None))
    assert np.linalg.norm(out - c["maskFT"]) / np.linalg.norm(c["maskFT"]) < H.TOL


def _emu_direct_operator(lib, pn, ps, sign):
    A = np.zeros((pn, pn), np.complex64)
    lib.check(lib.litho_direct_operator(pn, float(ps), 193.0, sign, A.ctypes.data, None))
    return A


@pytest.mark.parametrize("name", ["direct_64", "direct_128"])
def test_emu_direct_solver_matches_reference(name):
    """abbeImage(fft=False) and Mask.fraunhofer(fft=False) through the direct-solver kernels."""
    import ctypes as C
    c = KAT[name]
    lib = H.emu_lib()
    pn = c["maskFT"].shape[0]
    ps = int(c["pixel_size"])
    A = _emu_direct_operator(lib, pn, ps, -1)
    assert np.abs(A - O.direct_operator(pn, ps, 193.0, -1.0, np.complex128)).max() < 2e-6
    mft = np.ascontiguousarray(c["maskFT"])
    pup = np.ascontiguousarray(c["pupil"])
    bbox = lib.pupil_bbox(pup.ctypes.data, pn)
    box = (C.c_int * 4)(*bbox)
    shifts = np.ascontiguousarray(O.source_shifts(c["lightsource"], pn))
    n = len(shifts)
    nbytes = lib.litho_direct_workspace_bytes(pn, box, 2)
    ws = np.zeros(nbytes, np.uint8)
    img = np.zeros((pn, pn), np.float32)
    lib.check(lib.litho_direct_accumulate(A.ctypes.data, mft.ctypes.data, pup.ctypes.data, pn, box, shifts.ctypes.data,
                                          None, n, 2, img.ctypes.data, ws.ctypes.data, nbytes, None))
    assert O.rel_l2(img, c["image"]) < H.TOL
    # mask spectrum by the direct integral
    Ap = _emu_direct_operator(lib, pn, ps, +1)
    geom = np.ascontiguousarray(c["geometry"], dtype=np.int16)
    full = (C.c_int * 4)(0, pn - 1, 0, pn - 1)
    nbytes = lib.litho_direct_workspace_bytes(pn, full, 1)
    ws = np.zeros(nbytes, np.uint8)
    out = np.zeros((pn, pn), np.complex64)
    lib.check(lib.litho_direct_mask_spectrum(Ap.ctypes.data, geom.ctypes.data, pn, out.ctypes.data, ws.ctypes.data,
                                             nbytes, None))
    assert np.linalg.norm(out - c["maskFT"]) / np.linalg.norm(c["maskFT"]) < H.TOL


def test_emu_direct_field():
    import ctypes as C
    d = KAT["field_direct_64"]
    lib = H.emu_lib()
    A = _emu_direct_operator(lib, 64, 25, -1)
    pf = np.ascontiguousarray(d["pf"])
    mft = np.ascontiguousarray(d["maskFT"])
    box = (C.c_int * 4)(*lib.pupil_bbox(pf.ctypes.data, 64))
    nbytes = lib.litho_direct_workspace_bytes(64, box, 1)
    ws = np.zeros(nbytes, np.uint8)
    field = np.zeros((64, 64), np.complex64)
    lib.check(lib.litho_direct_field(A.ctypes.data, pf.ctypes.data, mft.ctypes.data, 64, box, field.ctypes.data,
                                     ws.ctypes.data, nbytes, None))
    assert np.linalg.norm(field - d["field"]) / np.linalg.norm(d["field"]) < H.TOL


def test_emu_builders_match_reference():
    """LightSource / Pupil builders (fp16 step replay) against the reference's own tensors and the oracle."""
    import ctypes as C
    import math
    from lithographysimulator_b200 import workloads as wl
    lib = H.emu_lib()
    c = KAT["demo64_quasar"]
    out = np.zeros((64, 64), np.int64)
    lib.check(lib.litho_source_build(64, 0.4, 0.8, 0.0, 0.0, 4, -math.pi / 8, out.ctypes.data, None))
    assert (out == c["lightsource"]).all()
    lib.check(lib.litho_source_build(64, 0.4, 0.8, 0.0, 0.0, 0, 0.0, out.ctypes.data, None))
    assert (out == KAT["demo64_annular"]["lightsource"]).all()
    out = np.zeros((128, 128), np.int64)
    lib.check(lib.litho_source_build(128, 0.4, 0.8, 0.5, -0.25, 4, -math.pi / 8, out.ctypes.data, None))
    assert (out * wl.lattice(128, 7) == KAT["shifted_128"]["lightsource"]).all()
    for pn, si, so, cnt, rot in ((256, 0.6, 0.9, 0, 0.0), (256, 0.3, 1.3, 3, 0.3), (512, 0.0, 0.6, 5, 1.0)):
        out = np.zeros((pn, pn), np.int64)
        lib.check(lib.litho_source_build(pn, si, so, 0.0, 0.0, cnt, rot, out.ctypes.data, None))
        ref = O.light_source_annular(si, so, pn) if cnt == 0 else O.light_source_quasar(si, so, pn, cnt, rot)
        assert (out == ref).all()
    # pupil: coefficients after the reference's in-place defocus rescale (fp16 values)
    _, ab = O.wavefront_error(wl.ABERR_FULL, 64, 0.7, 193.0)
    arr = (C.c_float * len(ab))(*[float(v) for v in ab])
    pup = np.zeros((64, 64), np.complex64)
    we = np.zeros((64, 64), np.complex64)
    lib.check(lib.litho_pupil_build(arr, len(ab), 64, pup.ctypes.data, we.ctypes.data, None))
    assert np.abs(pup - c["pupil"]).max() < 2e-7          # the reference's own pupil tensor
    we_ref, _ = O.wavefront_error(wl.ABERR_FULL, 64, 0.7, 193.0)
    assert (we.real == we_ref).all() and not we.imag.any()
    z = np.load(f"{H.GOLDEN}/cfg1.npz")
    _, ab = O.wavefront_error([0, 0, 0, 0, 50], 256, 0.7, 193.0)
    arr = (C.c_float * len(ab))(*[float(v) for v in ab])
    pup = np.zeros((256, 256), np.complex64)
    lib.check(lib.litho_pupil_build(arr, len(ab), 256, pup.ctypes.data, None, None))
    assert np.abs(pup - z["pupil"]).max() < 2e-7


def test_emu_rejects_unsupported():
    lib = H.emu_lib()
    with pytest.raises(Exception):
        lib.plan_create(63, 128, (0, 62, 0, 62))       # odd grid
    with pytest.raises(Exception):
        lib.plan_create(256, 128, (0, 255, 0, 255))    # N < pn: the reference raises too (Q7)
    with pytest.raises(Exception):
        lib.plan_create(64, 100, (0, 63, 0, 63))       # N not a power of two


@pytest.mark.parametrize("name,fast", [("demo64_quasar", True), ("wrap_128", False)])
@pytest.mark.parametrize("tma", ["1", "0"])
def test_emu_focus_batching_equals_single_images(monkeypatch, name, fast, tma):
    """litho_abbe_fft_accumulate_focus (f1: one row pass serves all focus values, one column pass per focus value)
    against n_focus separate litho_abbe_fft_accumulate calls: bit-identical planes, on the fast path (TMA-staged and
    plain column kernels, uneven batches so the T ring wraps) and on the generic fallback; status words stay clear."""
    monkeypatch.setenv("LITHO_TMA", tma)
    c = KAT[name]
    lib = H.emu_lib()
    pn = c["maskFT"].shape[0]
    _, N = lib.epsilon_n(4 / pn, float(c["pixel_size"]), 193.0)
    mft = np.ascontiguousarray(c["maskFT"], np.complex64)
    base = np.ascontiguousarray(c["pupil"], np.complex64)
    # focus variants of one pupil: same support, different phase (defocus ~ r^2)
    yy, xx = np.mgrid[0:pn, 0:pn]
    r2 = ((yy - pn / 2) ** 2 + (xx - pn / 2) ** 2) / (pn / 4) ** 2
    F = 3
    pupils = np.ascontiguousarray(np.stack([base * np.exp(1j * 0.7 * f * r2) for f in range(F)]).astype(np.complex64))
    shifts = np.ascontiguousarray(O.source_shifts(c["lightsource"], pn)[::3][:11], np.int32)
    w = np.linspace(0.5, 1.5, len(shifts)).astype(np.float32)
    support = lib.pupil_support(base.ctypes.data, pn)
    plan = lib.plan_create(pn, N, support, 0 if fast else 1)
    assert plan.path == (2 if fast else 1)
    elems = plan.intensity_elems
    stride = elems + 5
    for batch in (0, 2):
        wsb = plan.workspace_bytes_focus(batch, F)
        ws = np.zeros(max(wsb, 8), np.uint8)
        planes = np.zeros((F, stride), np.float32)
        plan.accumulate_focus(mft.ctypes.data, pupils.ctypes.data, F, pn * pn, shifts.ctypes.data, w.ctypes.data,
                              len(shifts), batch, planes.ctypes.data, stride, ws.ctypes.data, wsb)
        for f in range(F):
            single = np.zeros(elems, np.float32)
            wsb1 = plan.workspace_bytes(batch)
            ws1 = np.zeros(max(wsb1, 8), np.uint8)
            plan.accumulate(mft.ctypes.data, pupils[f].ctypes.data, shifts.ctypes.data, w.ctypes.data, len(shifts),
                            batch, single.ctypes.data, ws1.ctypes.data, wsb1)
            assert np.array_equal(planes[f, :elems], single), (batch, f)
            assert not planes[f, elems:].any()
    assert plan.status() == (0, 0)
    # and against the oracle for one focus value
    out = np.zeros((pn, pn), np.float32)
    fwb = plan.finalize_workspace_bytes()
    fws = np.zeros(max(fwb, 8), np.uint8)
    plan.unpermute(planes[1].ctypes.data, out.ctypes.data, fws.ctypes.data, fwb)
    ref = np.zeros((pn, pn))
    for (d0, d1), wi in zip(shifts, w):
        ref += wi * np.abs(O.calculate_fft_aerial(np.roll(pupils[1], (int(d0), int(d1)), (0, 1)), mft, pn, N)) ** 2
    assert O.rel_l2(out, ref) < H.TOL
    plan.close()


def test_emu_status_word_reports_clamped_shift():
    """A fast plan driven outside its no-wrap contract through the raw C ABI: memory-safe, and litho_plan_status
    reports it (read-and-clear)."""
    c = KAT["demo64_quasar"]
    lib = H.emu_lib()
    mft = np.ascontiguousarray(c["maskFT"], np.complex64)
    pup = np.ascontiguousarray(c["pupil"], np.complex64)
    _, N = lib.epsilon_n(4 / 64, 25.0, 193.0)
    plan = lib.plan_create(64, N, lib.pupil_support(pup.ctypes.data, 64))
    assert plan.path == 2 and plan.status() == (0, 0)
    bad = np.array([[0, 0], [40, 0]], np.int32)
    inten = np.zeros(plan.intensity_elems, np.float32)
    wsb = plan.workspace_bytes(2)
    ws = np.zeros(wsb, np.uint8)
    plan.accumulate(mft.ctypes.data, pup.ctypes.data, bad.ctypes.data, None, 2, 2, inten.ctypes.data, ws.ctypes.data, wsb)
    assert plan.status() == (1, 0)
    assert plan.status() == (0, 0)
    plan.close()


def test_emu_subfft_4096_multi_tile_tma_column_pass():
    """Sub-FFT 4096 (the cfg5 kernel shape: radix 32 x 32 x 4, TMA-staged column pass whose exchange buffer, tile and
    tables fill the 227 KB of the CTA, third-pass twiddles read from global memory) on a synthetic 2100-px grid with a
    2060-px dense window, N = 8192, two source points in ONE batch -- the tile of point 1 is copied while the FFT of
    point 0 uses the exchange buffer, so any overlap of the two regions shows.  Against the oracle."""
    pn, win, ps = 2100, 2060, 12.37
    rng = np.random.default_rng(5)
    lo = pn // 2 - win // 2
    pup = np.zeros((pn, pn), np.complex64)
    pup[lo:lo + win, lo:lo + win] = rng.standard_normal((win, win)) + 1j * rng.standard_normal((win, win))
    mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
    shifts = np.array([[3, -5], [-7, 11]], np.int32)
    img, info = H.emu_abbe_fft(mft, pup, None, ps, 193.0, shifts=shifts, postprocess=False, batch=2)
    assert info["path"] == 2 and info["M"] == 4096 and info["N"] == 8192 and info["column_tile"] == 2, info
    ref = np.zeros((pn, pn))
    for d0, d1 in shifts:
        ref += np.abs(O.calculate_fft_aerial(np.roll(pup, (int(d0), int(d1)), (0, 1)), mft, pn, 8192)) ** 2
    assert O.rel_l2(img, ref) < H.TOL


def test_emu_subfft_1024_folded_row_pass():
    """Sub-FFT 1024 (the cfg3 kernel shape) on a synthetic 1060-px grid with a dense 1027-px window (S = M + 3), N = 4096:
    the row pass folds the residue-1 pre-twiddle into per-register constants and a second inter-pass table
    (FastShape::ROW_FOLD).  Two source points in one batch, against the oracle."""
    pn, win = 1060, 1027
    ps = pn * 193.0 / (4 * 4096)
    rng = np.random.default_rng(11)
    lo = pn // 2 - win // 2
    pup = np.zeros((pn, pn), np.complex64)
    pup[lo:lo + win, lo:lo + win] = rng.standard_normal((win, win)) + 1j * rng.standard_normal((win, win))
    mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
    shifts = np.array([[3, -5], [-7, 11]], np.int32)
    img, info = H.emu_abbe_fft(mft, pup, None, ps, 193.0, shifts=shifts, postprocess=False, batch=2)
    assert info["path"] == 2 and info["M"] == 1024 and info["N"] == 4096, info
    ref = np.zeros((pn, pn))
    for d0, d1 in shifts:
        ref += np.abs(O.calculate_fft_aerial(np.roll(pup, (int(d0), int(d1)), (0, 1)), mft, pn, 4096)) ** 2
    assert O.rel_l2(img, ref) < H.TOL
