"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the
reference-shaped Python API / C ABI, against (a) goldens produced by the reference itself and
(b) the oracle on the same seeded inputs.  Tolerance: relative L2 <= 1e-5 on the aerial image
(BASELINE.json north star).  Nothing here reads /root/reference."""
import math

import numpy as np
import pytest
import torch

import helpers as H
from oracle import abbe_oracle as O
from lithographysimulator_b200 import workloads as wl

pytestmark = pytest.mark.gpu

KAT = H.load_kat()
FFT_CASES = ["demo64_quasar", "demo64_annular", "ps50_64", "ps12_64", "np2_96", "wrap_128", "shifted_128", "dense_64"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def L(dev):
    import lithographysimulator_b200 as L
    from lithographysimulator_b200 import _native
    lib = _native.device_lib()  # raises if liblitho_b200.so is missing: no fallback
    assert lib.litho_is_device_build() == 1
    return L


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _mask_stub(L, pn, ps, dev):
    return L.Mask(torch.zeros((pn, pn), dtype=torch.int16), ps, dev)


@pytest.mark.parametrize("name", FFT_CASES)
def test_abbe_fft_matches_reference_golden(L, dev, name):
    c = KAT[name]
    pn = c["maskFT"].shape[0]
    ps = int(c["pixel_size"])
    m = _mask_stub(L, pn, ps, dev)
    img = L.abbeImage(m, _t(c["maskFT"], dev), _t(c["pupil"], dev), _t(c["lightsource"], dev), ps, m.deltaK, 193.0,
                      True, dev)
    assert img.dtype == torch.float32 and tuple(img.shape) == c["image"].shape
    assert O.rel_l2(img.cpu().numpy(), c["image"]) < H.TOL


def test_abbe_fft_accepts_host_tensors(L, dev):
    """The reference lets every argument live anywhere; host tensors are uploaded inside the call."""
    c = KAT["demo64_quasar"]
    m = _mask_stub(L, 64, 25, dev)
    img = L.abbeImage(m, torch.from_numpy(c["maskFT"]), torch.from_numpy(c["pupil"]),
                      torch.from_numpy(c["lightsource"]), 25, m.deltaK, 193.0, True, dev)
    assert O.rel_l2(img.cpu().numpy(), c["image"]) < H.TOL


def test_fft_field_matches_reference_golden(L, dev):
    f = KAT["field_fft_64"]
    e = L.calculateFFTAerial(_t(f["pf"], dev), _t(f["maskFT"], dev), 64, 128).cpu().numpy()
    assert e.dtype == np.complex64
    assert np.linalg.norm(e - f["field"]) / np.linalg.norm(f["field"]) < H.TOL


def test_cpu_device_is_rejected(L):
    c = KAT["demo64_quasar"]
    with pytest.raises(Exception):
        L.abbeImage(None, torch.from_numpy(c["maskFT"]), torch.from_numpy(c["pupil"]),
                    torch.from_numpy(c["lightsource"]), 25, 4 / 64, 193.0, True, torch.device("cpu"))


def test_empty_source_and_empty_pupil(L, dev):
    c = KAT["demo64_quasar"]
    m = _mask_stub(L, 64, 25, dev)
    z = L.abbeImage(m, _t(c["maskFT"], dev), _t(c["pupil"], dev), torch.zeros((64, 64), dtype=torch.int64, device=dev),
                    25, m.deltaK, 193.0, True, dev)
    assert tuple(z.shape) == (64, 64) and float(z.abs().max()) == 0.0
    z = L.abbeImage(m, _t(c["maskFT"], dev), torch.zeros((64, 64), dtype=torch.complex64, device=dev),
                    _t(c["lightsource"], dev), 25, m.deltaK, 193.0, True, dev)
    assert float(z.abs().max()) == 0.0


def _cfg_inputs(name):
    """Inputs of a BASELINE config from the oracle builders (pinned against the reference in the CPU suite)."""
    cfg = wl.CONFIGS[name]
    geom = cfg.geometry()
    mft = O.fraunhofer(geom, cfg.pixel_size, cfg.wavelength, True, np.complex128).astype(np.complex64)
    if cfg.source == "quasar":
        ls = O.light_source_quasar(cfg.sigma_in, cfg.sigma_out, cfg.pn, 4, -math.pi / 8)
    else:
        ls = O.light_source_annular(cfg.sigma_in, cfg.sigma_out, cfg.pn)
    ls = ls * wl.lattice(cfg.pn, cfg.stride)
    pf, _ = O.pupil_function(cfg.aberrations, cfg.pn, cfg.na, cfg.wavelength)
    return cfg, mft, pf, ls


def test_cfg1_full_image(L, dev, golden_dir):
    z = np.load(f"{golden_dir}/cfg1.npz")
    cfg = wl.CONFIGS["cfg1"]
    m = _mask_stub(L, cfg.pn, cfg.pixel_size, dev)
    ls = O.light_source_annular(cfg.sigma_in, cfg.sigma_out, cfg.pn) * wl.lattice(cfg.pn, cfg.stride)
    img = L.abbeImage(m, _t(z["maskFT"], dev), _t(z["pupil"], dev), _t(ls, dev), cfg.pixel_size, m.deltaK,
                      cfg.wavelength, True, dev).cpu().numpy()
    assert O.rel_l2(img, z["image"]) < H.TOL
    # float64 reference run (north star): same inputs through the oracle in complex128
    ref64 = O.abbe_image(z["maskFT"], z["pupil"], ls, cfg.pixel_size, 4 / cfg.pn, cfg.wavelength, True, np.complex128)
    assert O.rel_l2(img, ref64) < H.TOL


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_large_config_against_reference_sample(L, dev, golden_dir, name):
    """Full-size BASELINE configs: every 16th pixel of the reference's own image plus its moments."""
    z = np.load(f"{golden_dir}/{name}.npz")
    cfg, mft, pf, ls = _cfg_inputs(name)
    assert int((ls != 0).sum()) == int(z["n_src"])
    # the oracle-built inputs agree with the reference-built ones at the sampled positions
    st = int(z["sample_stride"])
    assert np.linalg.norm(mft[::st, ::st] - z["maskFT_sample"]) / np.linalg.norm(z["maskFT_sample"]) < 1e-6
    assert np.abs(pf[::st, ::st] - z["pupil_sample"]).max() < 1e-3
    m = _mask_stub(L, cfg.pn, cfg.pixel_size, dev)
    img = L.abbeImage(m, _t(mft, dev), _t(pf, dev), _t(ls, dev), cfg.pixel_size, m.deltaK, cfg.wavelength, True, dev)
    img = img.cpu().numpy()
    assert tuple(img.shape) == tuple(z["shape"])
    assert O.rel_l2(img[::st, ::st], z["image_sample"]) < H.TOL
    assert abs(img.sum(dtype=np.float64) / float(z["img_sum"]) - 1) < 1e-5
    assert abs((img.astype(np.float64) ** 2).sum() / float(z["img_sumsq"]) - 1) < 2e-5


def test_source_additivity_and_weights_full_size(L, dev):
    """Size-independent properties at the headline grid (2048 px): I(S1 u S2) = I(S1) + I(S2) -- what licenses
    sharding source points across GPUs -- and integer weights == repeated source points."""
    from lithographysimulator_b200.imaging import AbbeEngine
    cfg, mft, pf, ls = _cfg_inputs("cfg3")
    eng = AbbeEngine.get(dev)
    sh = torch.from_numpy(O.source_shifts(ls, cfg.pn))[:48]
    mft_d, pf_d = _t(mft, dev), _t(pf, dev)
    kw = dict(pixelSize=cfg.pixel_size, deltaK=4 / cfg.pn, wavelength=cfg.wavelength, postprocess=False)
    full = eng.abbe_fft(mft_d, pf_d, None, shifts=sh, **kw)
    a = eng.abbe_fft(mft_d, pf_d, None, shifts=sh[0::2], **kw)
    b = eng.abbe_fft(mft_d, pf_d, None, shifts=sh[1::2], batch=5, **kw)
    assert O.rel_l2((a + b).cpu().numpy(), full.cpu().numpy()) < 1e-6
    w = torch.ones(48)
    w[::3] = 2.0
    weighted = eng.abbe_fft(mft_d, pf_d, None, shifts=sh, weights=w, **kw)
    rep = eng.abbe_fft(mft_d, pf_d, None, shifts=torch.cat([sh, sh[::3]]), **kw)
    assert O.rel_l2(weighted.cpu().numpy(), rep.cpu().numpy()) < 1e-6
    # and one source point against the oracle's float64 field at full size
    one = eng.abbe_fft(mft_d, pf_d, None, shifts=sh[7:8], **kw).cpu().numpy()
    d0, d1 = (int(v) for v in sh[7])
    e = O.calculate_fft_aerial(np.roll(pf, (d0, d1), (0, 1)), mft, cfg.pn, 4096)
    assert O.rel_l2(one, np.abs(e) ** 2) < H.TOL


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("pn,box", [(512, (100, 400)), (1024, (256, 768)), (1024, (0, 1023))])
def test_subfft_sizes_dense_windows(L, dev, pn, box, generic):
    """Random dense windows that force sub-FFT lengths 512 / 512 / 1024, through the fast coarse-grid
    kernels (q = 1, q = 2 with fully populated rim lines, q = 1 with pn < N) and the generic ones."""
    from lithographysimulator_b200.imaging import AbbeEngine
    rng = np.random.default_rng(pn + box[0])
    lo, hi = box
    n = hi - lo + 1
    pup = np.zeros((pn, pn), np.complex64)
    pup[lo:hi + 1, lo:hi + 1] = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    mft = (rng.standard_normal((pn, pn)) + 1j * rng.standard_normal((pn, pn))).astype(np.complex64)
    sh = torch.tensor([[7, -11], [-3, 2]], dtype=torch.int32)
    eng = AbbeEngine.get(dev)
    img = eng.abbe_fft(_t(mft, dev), _t(pup, dev), None, 25, 4 / pn, 193.0, shifts=sh, postprocess=False,
                       generic=generic).cpu().numpy()
    ref = np.zeros((pn, pn))
    for d0, d1 in sh.numpy():
        ref += np.abs(O.calculate_fft_aerial(np.roll(pup, (d0, d1), (0, 1)), mft, pn, 2 * pn)) ** 2
    assert O.rel_l2(img, ref) < H.TOL


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_fast_path_equals_generic_path(L, dev, name):
    """The coarse-grid fast path (incl. rim lines and spectral interpolation) against the literal
    fine-grid kernels on the same inputs at the BASELINE sizes."""
    from lithographysimulator_b200.imaging import AbbeEngine
    cfg, mft, pf, ls = _cfg_inputs(name)
    eng = AbbeEngine.get(dev)
    sh = torch.from_numpy(O.source_shifts(ls, cfg.pn))[::41]
    kw = dict(pixelSize=cfg.pixel_size, deltaK=4 / cfg.pn, wavelength=cfg.wavelength, shifts=sh)
    mft_d, pf_d = _t(mft, dev), _t(pf, dev)
    fast = eng.abbe_fft(mft_d, pf_d, None, **kw).cpu().numpy()
    gen = eng.abbe_fft(mft_d, pf_d, None, generic=True, **kw).cpu().numpy()
    assert fast.shape == gen.shape
    assert O.rel_l2(fast, gen) < H.TOL
    support = eng.pupil_support(pf_d)
    assert eng.plan_for(cfg.pn, 2 * cfg.pn, support, sh.to(dev)).path == 2


def test_builders_match_oracle(L, dev):
    cfg = wl.CONFIGS["cfg1"]
    src = L.LightSource(0.4, 0.8, 256, 0.7, 0, 0, dev)
    assert (src.generateQuasar(4, -math.pi / 8).cpu().numpy() == O.light_source_quasar(0.4, 0.8, 256, 4, -math.pi / 8)).all()
    assert (src.generateAnnular().cpu().numpy() == O.light_source_annular(0.4, 0.8, 256)).all()
    ab = torch.tensor(wl.ABERR_FULL, dtype=torch.float16, device=dev)
    pf = L.Pupil(256, 193.0, 0.7, ab, dev).generatePupilFunction().cpu().numpy()
    ref, ab_ref = O.pupil_function(wl.ABERR_FULL, 256, 0.7, 193.0)
    assert ((pf != 0) == (ref != 0)).all()
    # native fp16-step replay with double-precision transcendentals: same tensor as the reference's CPU run
    assert np.abs(pf - ref).max() < 2e-7
    z = np.load(f"{H.GOLDEN}/cfg3.npz")
    ab3 = torch.tensor(wl.ABERR_FULL, dtype=torch.float16, device=dev)
    pf3 = L.Pupil(2048, 193.0, 0.7, ab3, dev).generatePupilFunction()
    assert int((pf3 != 0).sum()) == int(z["pupil_nnz"])
    assert np.abs(pf3[::16, ::16].cpu().numpy() - z["pupil_sample"]).max() < 2e-7   # the reference's own pupil
    assert float(ab[4]) == float(ab_ref[4])  # in-place defocus rescale, like the reference (Q4)
    m = L.Mask(torch.from_numpy(cfg.geometry()), 25, dev)
    mft = m.fraunhofer(193.0, True).cpu().numpy()
    ref = O.fraunhofer(cfg.geometry(), 25, 193.0, True)
    assert np.linalg.norm(mft - ref) / np.linalg.norm(ref) < H.TOL


@pytest.mark.parametrize("name", ["direct_64", "direct_128"])
def test_direct_solver_matches_reference_golden(L, dev, name):
    """abbeImage(fft=False) and Mask.fraunhofer(fft=False): the 'torch Abbe path' parity of the north star."""
    c = KAT[name]
    pn = c["maskFT"].shape[0]
    ps = int(c["pixel_size"])
    m = L.Mask(torch.from_numpy(c["geometry"]), ps, dev)
    img = L.abbeImage(m, _t(c["maskFT"], dev), _t(c["pupil"], dev), _t(c["lightsource"], dev), ps, m.deltaK, 193.0,
                      False, dev)
    assert tuple(img.shape) == (pn, pn) and img.dtype == torch.float32
    assert O.rel_l2(img.cpu().numpy(), c["image"]) < H.TOL
    mft = m.fraunhofer(193.0, False).cpu().numpy()
    assert np.linalg.norm(mft - c["maskFT"]) / np.linalg.norm(c["maskFT"]) < H.TOL


def test_direct_field_matches_reference_golden(L, dev):
    d = KAT["field_direct_64"]
    e = L.calculateAerial(_t(d["pf"], dev), _t(d["maskFT"], dev), (-2 * 1j * math.pi) / 193.0, 64, 25, dev)
    assert np.linalg.norm(e.cpu().numpy() - d["field"]) / np.linalg.norm(d["field"]) < H.TOL


def test_direct_solver_256_against_oracle(L, dev):
    """cfg1 grid (256 px) where the reference itself needs ~126 GB: checked against the oracle's A G A^T."""
    z = np.load(f"{H.GOLDEN}/cfg1.npz")
    cfg = wl.CONFIGS["cfg1"]
    ls = np.zeros((256, 256), np.int64)
    for r, c in ((128, 128), (100, 140), (150, 90), (128, 70)):
        ls[r, c] = 1
    mftd = O.fraunhofer(cfg.geometry(), 25, 193.0, False, np.complex128).astype(np.complex64)
    m = L.Mask(torch.from_numpy(cfg.geometry()), 25, dev)
    img = L.abbeImage(m, _t(mftd, dev), _t(z["pupil"], dev), _t(ls, dev), 25, m.deltaK, 193.0, False, dev).cpu().numpy()
    ref = O.abbe_image(mftd, z["pupil"], ls, 25, 4 / 256, 193.0, False, np.complex128)
    assert O.rel_l2(img, ref) < H.TOL


def test_focus_sweep_matches_single_images(L, dev):
    """BASELINE cfg5 semantics at a small grid: one image per defocus value, same mask spectrum and source."""
    from lithographysimulator_b200.distributed import focus_sweep_sharded
    c = KAT["demo64_quasar"]
    m = _mask_stub(L, 64, 25, dev)
    pupils = []
    for d in (-150.0, 0.0, 150.0):
        ab = torch.tensor([0, 0, 0.01, 0, d, 0.01], dtype=torch.float16, device=dev)
        pupils.append(L.Pupil(64, 193.0, 0.7, ab, dev).generatePupilFunction())
    imgs = focus_sweep_sharded(m, _t(c["maskFT"], dev), pupils, _t(c["lightsource"], dev), 25, m.deltaK, 193.0, dev)
    assert len(imgs) == 3
    for img, pf in zip(imgs, pupils):
        ref = O.abbe_image(c["maskFT"], pf.cpu().numpy(), c["lightsource"], 25, 4 / 64, 193.0, True, np.complex128)
        assert O.rel_l2(img.cpu().numpy(), ref) < H.TOL


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_tma_staged_column_pass_equals_plain_loads(L, dev, name):
    """Column pass with the T tile staged in shared memory by the TMA engine (cp.async.bulk.tensor + mbarrier,
    the default for sub-FFT <= 1024) against the same pass with plain global loads (LITHO_TMA=0), and both
    against the oracle on a few source points.  Uneven batches make the 3-slot T ring wrap."""
    import os
    from lithographysimulator_b200 import _native
    cfg, mft, pf, ls = _cfg_inputs(name)
    lib = _native.device_lib()
    pn, N = cfg.pn, 2 * cfg.pn
    sh = torch.from_numpy(O.source_shifts(ls, pn))[::13][:23].contiguous().to(dev)
    w = torch.linspace(0.5, 2.0, sh.shape[0], device=dev)
    mft_d, pf_d = _t(mft, dev), _t(pf, dev)
    support = lib.pupil_support(pf_d.data_ptr(), pn, 0)
    outs = {}
    for tma in ("1", "0"):
        os.environ["LITHO_TMA"] = tma
        try:
            plan = lib.plan_create(pn, N, support)     # the toggle is read when a plan is created
        finally:
            os.environ.pop("LITHO_TMA", None)
        assert plan.path == 2 and (plan.column_tile() > 0) == (tma == "1")
        inten = torch.zeros(plan.intensity_elems, dtype=torch.float32, device=dev)
        wsb = plan.workspace_bytes(4)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        plan.accumulate(mft_d.data_ptr(), pf_d.data_ptr(), sh.data_ptr(), w.data_ptr(), sh.shape[0], 4,
                        inten.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream(dev).cuda_stream)
        out = torch.empty((pn, pn), dtype=torch.float32, device=dev)
        fwb = plan.finalize_workspace_bytes()
        fws = torch.empty(max(fwb, 16), dtype=torch.uint8, device=dev)
        plan.unpermute(inten.data_ptr(), out.data_ptr(), fws.data_ptr(), fwb, torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)
        outs[tma] = out.cpu().numpy()
        plan.close()
    assert O.rel_l2(outs["1"], outs["0"]) < 1e-6
    if pn <= 1024:   # the oracle takes seconds per source point beyond that; cfg3 is pinned by the golden sample test
        ref = np.zeros((pn, pn))
        for (d0, d1), wi in zip(sh.cpu().numpy(), w.cpu().numpy()):
            ref += wi * np.abs(O.calculate_fft_aerial(np.roll(pf, (int(d0), int(d1)), (0, 1)), mft, pn, N)) ** 2
        assert O.rel_l2(outs["1"], ref) < H.TOL


def test_prepare_run_pipeline_matches_single_call(L, dev):
    """AbbeEngine.prepare()/run() (inputs staged from pinned host memory on a copy stream, one image ahead)
    returns the image abbeImage() returns, also when two images with different sources are in flight."""
    from lithographysimulator_b200.imaging import AbbeEngine
    cfg, mft, pf, ls = _cfg_inputs("cfg2")
    eng = AbbeEngine.get(dev)
    ls2 = ls * wl.lattice(cfg.pn, 3 * cfg.stride)
    host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (mft, pf, ls, ls2)]
    args = (cfg.pixel_size, 4 / cfg.pn, cfg.wavelength)
    p0 = eng.prepare(host[0], host[1], host[2], *args, slot=0)
    p1 = eng.prepare(host[0], host[1], host[3], *args, slot=1)
    i0 = eng.run(p0)
    p2 = eng.prepare(host[0], host[1], host[3], *args, slot=0)   # refills set 0 behind run(p0)
    i1 = eng.run(p1)
    i2 = eng.run(p2)
    torch.cuda.synchronize(dev)
    m = _mask_stub(L, cfg.pn, cfg.pixel_size, dev)
    r0 = L.abbeImage(m, host[0], host[1], host[2], *args, True, dev)
    r1 = L.abbeImage(m, host[0], host[1], host[3], *args, True, dev)
    assert torch.equal(i0, r0) and torch.equal(i1, r1) and torch.equal(i2, r1)


def test_end_to_end_object_api(L, dev):
    """The reference demo (imageformation.py:99-119) through the object API, against its golden image."""
    c = KAT["demo64_quasar"]
    m = L.Mask(device=dev, pixelSize=25)
    mft = m.fraunhofer(193.0, True)
    src = L.LightSource(sigmaIn=0.4, sigmaOut=0.8, device=dev)
    ls = src.generateQuasar(4, -math.pi / (4 * 2))
    ab = torch.tensor(wl.ABERR_FULL, dtype=torch.float16, device=dev)
    pf = L.Pupil(m.pixelNumber, 193.0, src.NA, ab, device=dev).generatePupilFunction()
    img = L.abbeImage(m, mft, pf, ls, m.pixelSize, m.deltaK, 193.0, True, dev).cpu().numpy()
    assert O.rel_l2(img, c["image"]) < H.TOL


def _dist_worker(rank, world, port, out_path):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import lithographysimulator_b200 as L
        from lithographysimulator_b200.distributed import abbe_image_sharded
        z = np.load(f"{H.GOLDEN}/cfg1.npz")
        cfg = wl.CONFIGS["cfg1"]
        ls = O.light_source_annular(cfg.sigma_in, cfg.sigma_out, cfg.pn) * wl.lattice(cfg.pn, cfg.stride)
        m = L.Mask(torch.zeros((cfg.pn, cfg.pn), dtype=torch.int16), cfg.pixel_size, dev)
        img = abbe_image_sharded(m, torch.from_numpy(z["maskFT"]), torch.from_numpy(z["pupil"]), torch.from_numpy(ls),
                                 cfg.pixel_size, m.deltaK, cfg.wavelength, dev)
        # pipelined form: each rank uploads half of every input tensor, the halves are all-gathered over NVLink
        from lithographysimulator_b200.imaging import AbbeEngine
        eng = AbbeEngine.get(dev)
        up = dist.new_group(backend="nccl")
        host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (z["maskFT"], z["pupil"], ls)]
        prep = eng.prepare(host[0], host[1], host[2], cfg.pixel_size, m.deltaK, cfg.wavelength, shard=(rank, world),
                           upload_group=up)
        img2 = eng.run(prep, reduce_fn=lambda t: dist.all_reduce(t))
        torch.cuda.synchronize(dev)
        assert torch.equal(img, img2), "sharded upload + prepare/run differs from abbe_image_sharded"
        # the same staging with copy-engine peer copies instead of the NCCL all-gather (distributed.PeerStaging)
        from lithographysimulator_b200.distributed import PeerStaging

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        stg = PeerStaging(eng.lib, AbbeEngine.peer_staging_bytes(cfg.pn, host[2].dtype), rank, world, exchange)
        for use in range(3):        # slot 0, 1, 0: the third use re-fills a slot the peer has pulled from
            prep = eng.prepare(host[0], host[1], host[2], cfg.pixel_size, m.deltaK, cfg.wavelength, slot=use % 2,
                               shard=(rank, world), upload_peers=stg)
            img3 = eng.run(prep, reduce_fn=lambda t: dist.all_reduce(t))
            torch.cuda.synchronize(dev)
            assert torch.equal(img, img3), f"peer-staged upload differs (use {use})"
        dist.barrier()
        stg.close()
        # pipelined throughput mode: 5 images, rotating root, partial planes summed over peer memory (CUDA IPC
        # + NVLink loads) and, for comparison, by ncclReduce; every image a root produced must equal `img`
        from lithographysimulator_b200.distributed import ShardedPipeline, shard_shifts
        from lithographysimulator_b200.imaging import source_shifts, epsilon_n
        mft_d, pf_d = torch.from_numpy(z["maskFT"]).to(dev), torch.from_numpy(z["pupil"]).to(dev)
        sh_all = source_shifts(torch.from_numpy(ls).to(dev), cfg.pn)
        eps, N = epsilon_n(m.deltaK, cfg.pixel_size, cfg.wavelength)
        plan = eng.plan_for(cfg.pn, N, eng.pupil_support(pf_d), sh_all)
        mine = shard_shifts(sh_all, rank, world)
        for mode in ("peer", "nccl"):
            pipe = ShardedPipeline(eng, plan, eps, reduce=mode)
            assert pipe.reduce == mode
            for i in range(5):
                pipe.submit(mft_d, pf_d, mine, inputs_ready=True)
                if i % world == rank:
                    pipe.join()
                    torch.cuda.synchronize(dev)
                    got = pipe.last_image
                    err = float((got.double() - img.double()).norm() / img.double().norm())
                    assert err < 1e-6, (mode, i, err)
            pipe.join()
            torch.cuda.synchronize(dev)
            pipe.check()
            pipe.close()
        if rank == 0:
            np.save(out_path, img.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_image(tmp_path, golden_dir):
    """Source points sharded over 2 GPUs + NCCL sum-reduce == reference image (skipped on a 1-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "img.npy")
    mp.spawn(_dist_worker, args=(2, port, out_path), nprocs=2, join=True)
    z = np.load(f"{golden_dir}/cfg1.npz")
    assert O.rel_l2(np.load(out_path), z["image"]) < H.TOL


def test_cfg4_size_single_points_against_oracle(L, dev):
    """BASELINE cfg4 grid (4096 px, N = 8192, sub-FFT 2048 = radix 32x32x2): two source points against the
    oracle's float64 field, and the 4094-px output side of the reference's post-processing."""
    from lithographysimulator_b200.imaging import AbbeEngine
    cfg = wl.CONFIGS["cfg4"]
    pn = cfg.pn
    geom = cfg.geometry()
    mft = O.fraunhofer(geom, cfg.pixel_size, cfg.wavelength, True, np.complex128).astype(np.complex64)
    pf, _ = O.pupil_function(cfg.aberrations, pn, cfg.na, cfg.wavelength)
    eng = AbbeEngine.get(dev)
    sh = torch.tensor([[300, -700], [-911, 5]], dtype=torch.int32)
    mft_d, pf_d = _t(mft, dev), _t(pf, dev)
    raw = eng.abbe_fft(mft_d, pf_d, None, cfg.pixel_size, 4 / pn, cfg.wavelength, shifts=sh, postprocess=False)
    ref = np.zeros((pn, pn))
    for d0, d1 in sh.numpy():
        ref += np.abs(O.calculate_fft_aerial(np.roll(pf, (d0, d1), (0, 1)), mft, pn, 2 * pn)) ** 2
    assert O.rel_l2(raw.cpu().numpy(), ref) < H.TOL
    img = eng.abbe_fft(mft_d, pf_d, None, cfg.pixel_size, 4 / pn, cfg.wavelength, shifts=sh)
    assert tuple(img.shape) == (4094, 4094)
    eps, _ = O.calculate_epsilon_n(4 / pn, cfg.pixel_size, cfg.wavelength)
    assert O.rel_l2(img.cpu().numpy(), O.fft_postprocess(ref, pn, eps, dtype=np.float64)) < H.TOL
    # mask spectrum of this size through the native kernels (N = 8192 transform, sub-FFT 8192)
    m = L.Mask(torch.from_numpy(geom), cfg.pixel_size, dev)
    mine = m.fraunhofer(cfg.wavelength, True).cpu().numpy()
    assert np.linalg.norm(mine - mft) / np.linalg.norm(mft) < H.TOL


@pytest.mark.parametrize("trim", [False, True])
def test_cfg5_size_fast_equals_generic(L, dev, trim):
    """BASELINE cfg5 grid (8192 px, N = 16384, sub-FFT 4096 = radix 32x32x4): fast coarse-grid path against the
    generic fine-grid kernels on one source point (the oracle would need a 4 GB complex128 FFT here).
    trim=False: the reference's own pupil, whose support is 4099 = M+3 px wide -- the fast path folds the two
    extra rows/columns and takes the frequency lines M..M+2 from the generalised rim sums; trim=True: the even
    fit M+1."""
    from lithographysimulator_b200.imaging import AbbeEngine
    pn = 8192
    ab = torch.tensor([0, 0, 0.01, 0, -150, 0.01], dtype=torch.float16, device=dev)
    pf = L.Pupil(pn, 193.0, 0.7, ab, dev).generatePupilFunction()
    eng = AbbeEngine.get(dev)
    # At 8192 px the reference's fp16 grid collapses neighbouring coordinates near |x| = 1, so its pupil
    # support is 4099 px wide (one more than 2*pn/4+1 on each side).
    assert eng.pupil_support(pf)[:4] == (2047, 6145, 2047, 6145)
    if trim:
        pf[:2048] = 0
        pf[6145:] = 0
        pf[:, :2048] = 0
        pf[:, 6145:] = 0
    g = torch.Generator(device="cpu").manual_seed(3)
    mft = torch.complex(torch.randn((pn, pn), generator=g), torch.randn((pn, pn), generator=g)).to(dev)
    sh = torch.tensor([[1500, -900]], dtype=torch.int32)
    kw = dict(pixelSize=25, deltaK=4 / pn, wavelength=193.0, shifts=sh, postprocess=False)
    fast = eng.abbe_fft(mft, pf, None, **kw)
    gen = eng.abbe_fft(mft, pf, None, generic=True, **kw)
    plan = eng.plan_for(pn, 2 * pn, eng.pupil_support(pf), sh.to(dev))
    assert plan.path == 2 and plan.M == 4096
    num = torch.linalg.vector_norm((fast - gen).double())
    den = torch.linalg.vector_norm(gen.double())
    assert float(num / den) < H.TOL


def test_cfg5_reference_pupil_against_oracle(L, dev):
    """BASELINE cfg5 grid with the reference's own 8192-px pupil (support 4099 px = sub-FFT 4096 + 3: fast path
    with folded extra rows/columns and three rim lines per side, N = 16384): one source point against the
    oracle's FFT solver (complex64 here: a 16384^2 complex128 plane
    would need 4 GB per copy; the oracle's c64/c128 gap is 1e-7)."""
    from lithographysimulator_b200.imaging import AbbeEngine
    pn = 8192
    pf_ref, _ = O.pupil_function([0, 0, 0.01, 0, -150, 0.01], pn, 0.7, 193.0)
    ab = torch.tensor([0, 0, 0.01, 0, -150, 0.01], dtype=torch.float16, device=dev)
    pf = L.Pupil(pn, 193.0, 0.7, ab, dev).generatePupilFunction()
    assert np.abs(pf.cpu().numpy() - pf_ref).max() < 2e-7
    rng = np.random.default_rng(8)
    mft = (rng.standard_normal((pn, pn), dtype=np.float32) + 1j * rng.standard_normal((pn, pn), dtype=np.float32))
    mft = mft.astype(np.complex64)
    sh = torch.tensor([[1500, -900]], dtype=torch.int32)
    eng = AbbeEngine.get(dev)
    img = eng.abbe_fft(_t(mft, dev), pf, None, 25, 4 / pn, 193.0, shifts=sh, postprocess=False).cpu().numpy()
    e = O.calculate_fft_aerial(np.roll(pf_ref, (1500, -900), (0, 1)), mft, pn, 2 * pn, np.complex64)
    ref = e.real.astype(np.float64) ** 2 + e.imag.astype(np.float64) ** 2
    assert O.rel_l2(img, ref) < H.TOL


@pytest.mark.parametrize("name", ["cfg4", "cfg5"])
def test_largest_grids_against_reference_subset_golden(L, dev, golden_dir, name):
    """BASELINE cfg4 (4096 px, N = 8192) and cfg5 (8192 px, N = 16384, the reference's 4099-px-wide pupil) through
    the whole object chain -- Mask.fraunhofer, Pupil, abbeImage, all on the GPU builders and kernels -- against
    images the unmodified reference produced for a handful of those configs' source points
    (tests/golden/cfg{4,5}_subset.npz, oracle/make_golden.py; the full configs take the reference hours)."""
    z = np.load(f"{golden_dir}/{name}_subset.npz")
    cfg = wl.CONFIGS[name]
    pn, st = cfg.pn, int(z["sample_stride"])
    mask = L.Mask(torch.from_numpy(cfg.geometry()), cfg.pixel_size, dev)
    mft = mask.fraunhofer(cfg.wavelength, True)
    ref_m = torch.from_numpy(z["maskFT_sample"]).to(dev)
    assert float(torch.linalg.vector_norm(mft[::st, ::st] - ref_m) / torch.linalg.vector_norm(ref_m)) < 1e-5
    ab = torch.tensor(z["aberrations"], dtype=torch.float16, device=dev)
    pf = L.Pupil(pn, cfg.wavelength, cfg.na, ab, dev).generatePupilFunction()
    assert int((pf != 0).sum()) == int(z["pupil_nnz"])
    assert float((pf[::st, ::st] - torch.from_numpy(z["pupil_sample"]).to(dev)).abs().max()) < 1e-3
    rows = torch.from_numpy(z["ls_rows"].astype(np.int64)).to(dev)
    ls = torch.zeros((pn, pn), dtype=torch.int64, device=dev)
    ls[rows[:, 0], rows[:, 1]] = 1
    img = L.abbeImage(mask, mft, pf, ls, cfg.pixel_size, mask.deltaK, cfg.wavelength, True, dev)
    assert tuple(img.shape) == tuple(z["shape"])
    got = img[::st, ::st].cpu().numpy()
    assert O.rel_l2(got, z["image_sample"]) < H.TOL
    assert abs(float(img.sum(dtype=torch.float64)) / float(z["img_sum"]) - 1) < 1e-5
    assert abs(float((img.double() ** 2).sum()) / float(z["img_sumsq"]) - 1) < 2e-5


@pytest.mark.parametrize("name", ["cfg4", "cfg5"])
def test_full_cfg4_cfg5_against_reference_golden(L, dev, golden_dir, name):
    """The complete BASELINE cfg4 (4104 source points) and cfg5 (980 points, first focus value) images against the
    unmodified reference's (tests/golden/cfg4.npz, cfg5.npz), through the object chain on the GPU."""
    import os
    path = f"{golden_dir}/{name}.npz"
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    z = np.load(path)
    cfg = wl.CONFIGS[name]
    pn, st = cfg.pn, int(z["sample_stride"])
    mask = L.Mask(torch.from_numpy(cfg.geometry()), cfg.pixel_size, dev)
    mft = mask.fraunhofer(cfg.wavelength, True)
    src = L.LightSource(cfg.sigma_in, cfg.sigma_out, pn, cfg.na, 0, 0, dev)
    ls = src.generateQuasar(4, -math.pi / 8) if cfg.source == "quasar" else src.generateAnnular()
    ls = ls * torch.from_numpy(wl.lattice(pn, cfg.stride)).to(dev)
    assert (torch.argwhere(ls).cpu().numpy() == z["ls_rows"]).all()
    ab = torch.tensor(z["aberrations"], dtype=torch.float16, device=dev)
    pf = L.Pupil(pn, cfg.wavelength, cfg.na, ab, dev).generatePupilFunction()
    img = L.abbeImage(mask, mft, pf, ls, cfg.pixel_size, mask.deltaK, cfg.wavelength, True, dev)
    assert tuple(img.shape) == tuple(z["shape"])
    assert O.rel_l2(img[::st, ::st].cpu().numpy(), z["image_sample"]) < H.TOL
    assert abs(float(img.sum(dtype=torch.float64)) / float(z["img_sum"]) - 1) < 1e-5
    assert abs(float((img.double() ** 2).sum()) / float(z["img_sumsq"]) - 1) < 2e-5


def test_sharded_pipeline_single_gpu_chained(L, dev):
    """ShardedPipeline on one GPU (what bench.py times): 4 images queued back to back with inputs_ready=True, so
    the row pass of image i+1 is chained behind image i on the plan's auxiliary stream and the post-processing of
    image i overlaps the accumulation of image i+1.  Every image must be bit-identical to the plain call (the rim
    sums are deterministic) and match the reference golden."""
    from lithographysimulator_b200.distributed import ShardedPipeline
    from lithographysimulator_b200.imaging import AbbeEngine, source_shifts, epsilon_n
    cfg, mft, pf, ls = _cfg_inputs("cfg2")
    eng = AbbeEngine.get(dev)
    mft_d, pf_d, ls_d = _t(mft, dev), _t(pf, dev), _t(ls, dev)
    args = (cfg.pixel_size, 4 / cfg.pn, cfg.wavelength)
    ref = eng.abbe_fft(mft_d, pf_d, ls_d, *args)
    eps, N = epsilon_n(4 / cfg.pn, cfg.pixel_size, cfg.wavelength)
    sh = source_shifts(ls_d, cfg.pn)
    plan = eng.plan_for(cfg.pn, N, eng.pupil_support(pf_d), sh)
    assert plan.path == 2
    pipe = ShardedPipeline(eng, plan, eps)
    for i in range(4):
        pipe.submit(mft_d, pf_d, sh, inputs_ready=True, batch=(0 if i != 2 else 37))   # image 2 breaks the chain
    pipe.join()
    torch.cuda.synchronize(dev)
    assert torch.equal(pipe.last_image, ref)
    assert plan.status() == (0, 0)
    # images in flight use alternating output buffers: check the last two
    assert torch.equal(pipe.images[0], ref) and torch.equal(pipe.images[1], ref)


def test_chained_images_with_interleaved_plans_and_field_calls(L, dev):
    """chained accumulate calls (LITHO_PHASE_INPUTS_READY) interleaved with a second plan, an fft_field
    call and a direct-solver call on the same engine.  Each fast plan owns its T ring, so nothing the other calls
    queue can touch it; every chained image must equal its unchained twin."""
    from lithographysimulator_b200.imaging import AbbeEngine, source_shifts, epsilon_n
    cfg, mft, pf, ls = _cfg_inputs("cfg2")
    eng = AbbeEngine.get(dev)
    pn = cfg.pn
    mft_d, pf_d, ls_d = _t(mft, dev), _t(pf, dev), _t(ls, dev)
    eps, N = epsilon_n(4 / pn, cfg.pixel_size, cfg.wavelength)
    sh = source_shifts(ls_d, pn)
    planA = eng.plan_for(pn, N, eng.pupil_support(pf_d), sh)
    # a second fast plan: same pupil cropped to a smaller window
    pf2 = pf.copy()
    pf2[: pn // 2 - 100] = 0
    pf2[pn // 2 + 100:] = 0
    pf2_d = _t(pf2, dev)
    planB = eng.plan_for(pn, N, eng.pupil_support(pf2_d), sh)
    assert planA.path == 2 and planB.path == 2 and planA.handle.value != planB.handle.value
    c64 = KAT["demo64_quasar"]

    def unchained(plan, pupil):
        inten = eng.intensity_plane(plan)
        eng.accumulate(plan, mft_d, pupil, sh, inten)
        return inten.clone()

    refA, refB = unchained(planA, pf_d), unchained(planB, pf2_d)
    torch.cuda.synchronize(dev)
    outs = []
    for i in range(3):
        ia, ib = eng.intensity_plane(planA), eng.intensity_plane(planB)
        eng.accumulate(planA, mft_d, pf_d, sh, ia, None, 0 if i != 1 else 29, inputs_ready=True)
        L.calculateFFTAerial(_t(c64["pupil"], dev), _t(c64["maskFT"], dev), 64, 128)   # shares the engine
        eng.accumulate(planB, mft_d, pf2_d, sh, ib, None, 0, inputs_ready=True)
        outs.append((ia, ib))
    torch.cuda.synchronize(dev)
    for ia, ib in outs:
        assert torch.equal(ia, refA)
        assert torch.equal(ib, refB)
    assert planA.status() == (0, 0) and planB.status() == (0, 0)


def test_input_validation_and_error_words(L, dev):
    """ADVICE r1 / VERDICT r1 weak-8: mismatched shapes raise instead of indexing out of bounds; a pinned fast plan
    refuses sources that wrap the pupil window; and a raw C-ABI call that violates the no-wrap contract raises the
    plan's sticky error word instead of returning a silently wrong image."""
    from lithographysimulator_b200 import _native
    from lithographysimulator_b200.imaging import AbbeEngine
    c = KAT["demo64_quasar"]
    eng = AbbeEngine.get(dev)
    m = _mask_stub(L, 64, 25, dev)
    mft, pf, ls = _t(c["maskFT"], dev), _t(c["pupil"], dev), _t(c["lightsource"], dev)
    with pytest.raises(_native.LithoError):
        L.abbeImage(m, mft, pf[:32, :32].contiguous(), ls, 25, m.deltaK, 193.0, True, dev)     # smaller pupil
    with pytest.raises(_native.LithoError):
        L.abbeImage(m, mft, pf[:, :32].contiguous(), ls, 25, m.deltaK, 193.0, True, dev)       # non-square pupil
    with pytest.raises(_native.LithoError):
        L.abbeImage(m, mft, pf, ls[:32, :32].contiguous(), 25, m.deltaK, 193.0, True, dev)     # smaller source
    with pytest.raises(_native.LithoError):
        eng.abbe_fft(mft, pf, ls, 25, m.deltaK, 193.0, weights=torch.ones(3, device=dev))       # weight count
    with pytest.raises(_native.LithoError):
        L.calculateFFTAerial(pf[:32, :32].contiguous(), mft, 64, 128)
    # pinned fast plan + a source point that wraps the window
    eps, N = eng.lib.epsilon_n(m.deltaK, 25, 193.0)
    plan = eng.plan(64, N, eng.pupil_support(pf))
    assert plan.path == 2
    bad = torch.tensor([[0, 0], [40, 0]], dtype=torch.int32, device=dev)
    with pytest.raises(_native.LithoError):
        eng.abbe_fft(mft, pf, None, 25, m.deltaK, 193.0, shifts=bad, plan=plan)
    plan64 = eng.plan(64, N, eng.pupil_support(pf))
    with pytest.raises(_native.LithoError):
        eng.abbe_fft(_t(KAT["wrap_128"]["maskFT"], dev), _t(KAT["wrap_128"]["pupil"], dev), None, 25, 4 / 128, 193.0,
                     shifts=bad, plan=plan64)                                                   # plan of another grid
    # raw C ABI, contract violated: memory-safe, and reported by litho_plan_status
    assert plan.status() == (0, 0)
    inten = eng.intensity_plane(plan)
    wsb = plan.workspace_bytes(2)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    plan.accumulate(mft.data_ptr(), pf.data_ptr(), bad.data_ptr(), None, 2, 2, inten.data_ptr(), ws.data_ptr(), wsb,
                    torch.cuda.current_stream(dev).cuda_stream)
    assert plan.status() == (1, 0)
    assert plan.status() == (0, 0)      # read-and-clear


def test_builders_match_reference_on_this_gpu(L, dev):
    """the native LightSource / Pupil / Mask.fraunhofer builders against the UNMODIFIED reference
    (oracle/_ref, staged by oracle/build_ref.py) run with device='cuda' on this very GPU -- random rotation, count,
    shift, sigma, NA and aberrations, including non-power-of-two grids.  Sources and pupil supports must be
    identical, mask spectra within float32 rounding, pupil phases within two fp16 ulps of the wavefront error on at
    most a few percent of the pixels (CPU-vs-CUDA ATen rounding of r**3, see below)."""
    from oracle import ref_runner as RR
    if not RR.available():
        pytest.skip("oracle/_ref not staged")
    R = RR.load()
    rng = np.random.default_rng(20261017)
    bad = []
    for case in range(40):
        pn = int(rng.choice([64, 96, 128, 200, 256, 512]))
        s_in = float(rng.uniform(0.0, 0.6))
        s_out = float(s_in + rng.uniform(0.1, 0.39))
        count = int(rng.integers(1, 9))
        rot = float(rng.uniform(-math.pi, math.pi))
        sx, sy = (float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.3))) if case % 3 == 0 else (0, 0)
        ours = L.LightSource(s_in, s_out, pn, 0.7, sx, sy, dev).generateQuasar(count, rot)
        ref = R["lightsource"].LightSource(s_in, s_out, pn, 0.7, sx, sy, dev).generateQuasar(count, rot)
        d = int((ours != ref).sum())
        ours_a = L.LightSource(s_in, s_out, pn, 0.7, sx, sy, dev).generateAnnular()
        ref_a = R["lightsource"].LightSource(s_in, s_out, pn, 0.7, sx, sy, dev).generateAnnular()
        da = int((ours_a != ref_a).sum())
        if d or da:
            bad.append(("source", case, pn, count, rot, sx, sy, d, da))
        na = float(rng.uniform(0.5, 0.9))
        ab = (rng.uniform(-0.05, 0.05, int(rng.integers(5, 13)))).astype(np.float32)
        ab[4] = float(rng.uniform(-150, 150))
        p_ours = L.Pupil(pn, 193.0, na, torch.tensor(ab, dtype=torch.float16, device=dev), dev).generatePupilFunction()
        p_ref = R["pupil"].Pupil(pn, 193.0, na, torch.tensor(ab, dtype=torch.float16, device=dev), dev).generatePupilFunction()
        nz = int(((p_ours != 0) != (p_ref != 0)).sum())
        dp = float((p_ours - p_ref).abs().max())
        frac = float(((p_ours - p_ref).abs() > 1e-6).sum()) / (pn * pn)
        # Support identical.  Values: the native replay models the reference's CPU tensors (the goldens are CPU
        # runs); CUDA ATen rounds r**3 twice where CPU ATen rounds once, which moves the fp16 wavefront error of a
        # few pixels by one or two ulps (2^-11 .. 2^-10 waves = 3.1e-3 .. 6.2e-3 rad) -- nothing larger may appear.
        if nz or dp > 2 * math.pi * 2 ** -9 or frac > 0.05:
            bad.append(("pupil", case, pn, na, nz, dp, frac))
    assert not bad, bad
    for pn, ps in ((64, 25), (96, 25), (128, 50), (128, 12), (256, 25)):
        g = torch.from_numpy(wl.manhattan(pn, seed=pn))
        ours = L.Mask(g, ps, dev).fraunhofer(193.0, True)
        ref = R["mask"].Mask(g.to(dev), ps, dev).fraunhofer(193.0, True)
        assert float((ours - ref).abs().max() / ref.abs().max()) < 2e-6, (pn, ps)


def test_focus_batching_16_values_equals_single_images(L, dev):
    """f1 (BASELINE cfg5 semantics at 256 px): 16 defocus values through AbbeEngine.abbe_fft_focus -- one row pass
    serves all focus values, one column pass each -- against 16 single abbeImage calls (bit-identical) and the
    oracle's float64 image for three of them; also in groups of 5 (ragged last group)."""
    from lithographysimulator_b200.imaging import AbbeEngine
    cfg, mft, _, ls = _cfg_inputs("cfg1")
    eng = AbbeEngine.get(dev)
    pn = cfg.pn
    mft_d, ls_d = _t(mft, dev), _t(ls, dev)
    m = _mask_stub(L, pn, cfg.pixel_size, dev)
    pupils = []
    for d in np.linspace(-150, 150, 16):
        ab = torch.tensor([0, 0, 0.01, 0, float(d), 0.01, 0, 0.01, 0.01, 0.01], dtype=torch.float16, device=dev)
        pupils.append(L.Pupil(pn, cfg.wavelength, cfg.na, ab, dev).generatePupilFunction())
    args = (cfg.pixel_size, 4 / pn, cfg.wavelength)
    together = eng.abbe_fft_focus(mft_d, pupils, ls_d, *args)
    grouped = eng.abbe_fft_focus(mft_d, torch.stack(pupils), ls_d, *args, focus_batch=5)
    assert len(together) == 16 and len(grouped) == 16
    for f, pf in enumerate(pupils):
        single = L.abbeImage(m, mft_d, pf, ls_d, *args, True, dev)
        assert torch.equal(together[f], single), f
        assert torch.equal(grouped[f], single), f
        if f in (0, 7, 15):
            ref = O.abbe_image(mft, pf.cpu().numpy(), ls, cfg.pixel_size, 4 / pn, cfg.wavelength, True, np.complex128)
            assert O.rel_l2(together[f].cpu().numpy(), ref) < H.TOL
    # pupils with different supports are refused
    small = pupils[0].clone()
    small[: pn // 2 - 10] = 0
    from lithographysimulator_b200 import _native
    with pytest.raises(_native.LithoError):
        eng.abbe_fft_focus(mft_d, [pupils[0], small], ls_d, *args)


def test_cfg5_focus_pair_against_reference_golden(L, dev, golden_dir):
    """The first two focus values of the BASELINE cfg5 sweep (8192 px, 980 source points, sub-FFT 4096) imaged
    together through the focus-batched path; focus 0 against the unmodified reference's full image
    (tests/golden/cfg5.npz), focus 1 against the single-image call."""
    import os
    from lithographysimulator_b200.imaging import AbbeEngine
    path = f"{golden_dir}/cfg5.npz"
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    z = np.load(path)
    cfg = wl.CONFIGS["cfg5"]
    pn, st = cfg.pn, int(z["sample_stride"])
    eng = AbbeEngine.get(dev)
    mask = L.Mask(torch.from_numpy(cfg.geometry()), cfg.pixel_size, dev)
    mft = mask.fraunhofer(cfg.wavelength, True)
    src = L.LightSource(cfg.sigma_in, cfg.sigma_out, pn, cfg.na, 0, 0, dev)
    ls = src.generateQuasar(4, -math.pi / 8) * torch.from_numpy(wl.lattice(pn, cfg.stride)).to(dev)
    pupils = []
    for d in cfg.defocus_sweep[:2]:
        ab = list(z["aberrations"])
        ab[4] = d
        pupils.append(L.Pupil(pn, cfg.wavelength, cfg.na, torch.tensor(ab, dtype=torch.float16, device=dev), dev)
                      .generatePupilFunction())
    assert float(z["aberrations"][4]) == cfg.defocus_sweep[0]
    imgs = eng.abbe_fft_focus(mft, pupils, ls, cfg.pixel_size, mask.deltaK, cfg.wavelength)
    assert tuple(imgs[0].shape) == tuple(z["shape"])
    assert O.rel_l2(imgs[0][::st, ::st].cpu().numpy(), z["image_sample"]) < H.TOL
    assert abs(float(imgs[0].sum(dtype=torch.float64)) / float(z["img_sum"]) - 1) < 1e-5
    single = L.abbeImage(mask, mft, pupils[1], ls, cfg.pixel_size, mask.deltaK, cfg.wavelength, True, dev)
    assert torch.equal(imgs[1], single)


@pytest.mark.parametrize("name", ["cfg4", "cfg5"])
def test_tma_column_pass_large_subfft_multi_tile(L, dev, name):
    """Sub-FFT 2048 / 4096 (cfg4 / cfg5 grids): the TMA-staged column pass with several source points per launch
    (tile of point sl+1 copied while the FFT of sl runs; at 4096 the exchange buffer, the tile and the tables fill
    the 227 KB of the CTA and the third-pass twiddles come from global memory) against the plain-load kernel, with
    batches of 2 so that the 3-slot T ring wraps."""
    import os
    from lithographysimulator_b200 import _native
    cfg = wl.CONFIGS[name]
    lib = _native.device_lib()
    pn, N = cfg.pn, 2 * cfg.pn
    mft_d = L.Mask(torch.from_numpy(cfg.geometry()), cfg.pixel_size, dev).fraunhofer(cfg.wavelength, True)
    ab = torch.tensor(wl.aberrations_of(cfg), dtype=torch.float16, device=dev)
    pf_d = L.Pupil(pn, cfg.wavelength, cfg.na, ab, dev).generatePupilFunction()
    sh = torch.tensor([[0, 0], [pn // 5, -pn // 7], [-pn // 6, pn // 9], [17, 3], [-pn // 8, -pn // 8], [5, pn // 5],
                       [-3, 1]], dtype=torch.int32, device=dev)
    w = torch.linspace(0.5, 2.0, sh.shape[0], device=dev)
    support = lib.pupil_support(pf_d.data_ptr(), pn, 0)
    outs = {}
    for tma in ("1", "0"):
        os.environ["LITHO_TMA"] = tma
        try:
            plan = lib.plan_create(pn, N, support)
        finally:
            os.environ.pop("LITHO_TMA", None)
        assert plan.path == 2 and (plan.column_tile() > 0) == (tma == "1"), (plan.path, plan.column_tile())
        inten = torch.zeros(plan.intensity_elems, dtype=torch.float32, device=dev)
        wsb = plan.workspace_bytes(2)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        plan.accumulate(mft_d.data_ptr(), pf_d.data_ptr(), sh.data_ptr(), w.data_ptr(), sh.shape[0], 2,
                        inten.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)
        assert plan.status() == (0, 0)
        outs[tma] = inten.clone()
        plan.close()
        del ws
    a, b = outs["1"].double(), outs["0"].double()
    assert float((a - b).norm() / b.norm()) < 1e-6


@pytest.mark.parametrize("pn", [64, 128, 200, 256])
def test_direct_solver_tensor_core_path(L, dev, pn):
    """f2: the tcgen05 3xTF32 form of the direct solver (csrc/direct_tc.cu, default for pn >= 64) against the FP32
    CUDA-core kernels (LITHO_DIRECT_TC=0) and a float64 evaluation of the same operator, with weights, a shifted
    source whose rolled pupil wraps around the grid, and a grid that is not a multiple of the 128-row tile."""
    import ctypes as C
    import os
    from lithographysimulator_b200 import direct as D
    from lithographysimulator_b200.imaging import AbbeEngine, source_shifts
    eng = AbbeEngine.get(dev)
    geom = torch.from_numpy(wl.manhattan(pn, seed=7 + pn))
    mask = L.Mask(geom, 25, dev)
    mft = mask.fraunhofer(193.0, False)
    src = L.LightSource(0.5, 0.95, pn, 0.7, 0.2, -0.1, dev)           # shifted: some source points wrap the window
    stride = max(1, pn // 16)
    ls = src.generateQuasar(4, 0.3) * torch.from_numpy(wl.lattice(pn, stride)).to(dev)
    ab = torch.tensor([0, 0, 0.02, 0, 60, 0.01], dtype=torch.float16, device=dev)
    pf = L.Pupil(pn, 193.0, 0.7, ab, dev).generatePupilFunction()
    n = int((ls != 0).sum())
    assert n >= 8
    w = torch.linspace(0.5, 2.0, n, device=dev)
    outs = {}
    for tc in ("1", "0"):
        os.environ["LITHO_DIRECT_TC"] = tc
        try:
            outs[tc] = D.direct_abbe_image(mft, pf, ls, 25, 193.0, dev, weights=w, batch=5).double()
            torch.cuda.synchronize(dev)
        finally:
            os.environ.pop("LITHO_DIRECT_TC", None)
    st = C.c_int(0)
    eng.lib.check(eng.lib.litho_direct_status(C.byref(st), 0), "litho_direct_status")
    assert st.value == 0
    # float64 evaluation of E = A G A^T with the library's own operator
    A = D._operator(eng.lib, pn, 25, 193.0, -1, dev).to(torch.complex128)
    sh = source_shifts(ls, pn).cpu().numpy()
    ref = torch.zeros((pn, pn), dtype=torch.float64, device=dev)
    for (d0, d1), wi in zip(sh, w.cpu().numpy()):
        G = (torch.roll(pf, (int(d0), int(d1)), (0, 1)) * mft).to(torch.complex128)
        E = A @ G @ A.T
        ref += float(wi) * (E.real ** 2 + E.imag ** 2)
    for tc in ("1", "0"):
        err = float((outs[tc] - ref).norm() / ref.norm())
        assert err < 1e-5, (tc, err)
    assert float((outs["1"] - outs["0"]).norm() / outs["0"].norm()) < 1e-5


def test_source_points_kernel_matches_argwhere(L, dev):
    """litho_source_points (count + compact launches: ordered compaction of the non-zero source pixels, interleaved shard, shift
    bounds) against torch.argwhere for sparse, dense and empty planes, every element size, -0.0, and shards."""
    from lithographysimulator_b200 import _native
    lib = _native.device_lib()
    g = torch.Generator().manual_seed(3)
    cases = []
    for pn, density in ((64, 0.3), (250, 0.001), (256, 1.0), (1024, 0.0005), (2048, 0.00025), (128, 0.0)):
        base = (torch.rand((pn, pn), generator=g) < density)
        for dt in (torch.int64, torch.int32, torch.int16, torch.uint8, torch.bool, torch.float32, torch.float16, torch.float64):
            t = base.to(dt)
            if dt.is_floating_point:
                t = t * 2.5
                t[0, 0] = -0.0           # negative zero is not a source point
            cases.append((pn, t))
    for pn, t in cases[:: 3] + cases[1:: 7]:
        t_d = t.to(dev)
        ref = (torch.argwhere(t_d) - pn // 2).to(torch.int32)
        n = int(ref.shape[0])
        for rank, world in ((0, 1), (1, 3), (7, 8)):
            cap = max(1, n)
            out = torch.full((cap, 2), -12345, dtype=torch.int32, device=dev)
            n_all, n_mine, bounds = lib.source_points(t_d.data_ptr(), t_d.element_size(), t_d.is_floating_point(), pn, rank,
                                                      world, out.data_ptr(), cap, torch.cuda.current_stream(dev).cuda_stream)
            mine = ref[rank::world]
            assert n_all == n and n_mine == int(mine.shape[0]), (pn, t.dtype, rank, world, n_all, n, n_mine)
            assert torch.equal(out[:n_mine], mine), (pn, t.dtype, rank, world)
            if n:
                assert bounds == (int(ref[:, 0].min()), int(ref[:, 0].max()), int(ref[:, 1].min()), int(ref[:, 1].max()))
            else:
                assert bounds == (0, 0, 0, 0)
