"""Image formation: the reference's ``imageformation`` call surface on the B200-native kernels.

Mirrors, by name and positional order:

    abbeImage(mask, maskFT, pupilF, lightsource, pixelSize, deltaK, wavelength, fft, device)
    calculateFFTAerial(pf, maskFFFT, pixelNumber, N)
    calculateAerial(pupil, maskFT, fraunhoferConstant, pixelNumber, pixelSize, device)

(reference imageformation.py:47, :32, :3).  The arithmetic runs in hand-written sm_100a CUDA
behind the C ABI of include/litho_b200.h; torch is used for device memory and streams only.
There is no CPU path: a non-CUDA device raises.
"""
from __future__ import annotations

import threading

import torch

from . import _native

__all__ = ["abbeImage", "calculateFFTAerial", "calculateAerial", "AbbeEngine", "PreparedImage", "source_shifts",
           "epsilon_n", "tensor_from_ptr"]


def _require_cuda(device) -> torch.device:
    dev = torch.device(device) if device is not None else torch.device("cuda")
    if dev.type != "cuda":
        raise _native.LithoError(
            f"lithographysimulator_b200 runs on CUDA devices only (got device={dev}); there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def epsilon_n(deltaK: float, pixelSize, wavelength: float):
    """Mask.calculateEpsilonN (reference mask.py:63-72) evaluated on the host."""
    return _native.device_lib().epsilon_n(float(deltaK), float(pixelSize), float(wavelength))


def source_shifts(lightsource: torch.Tensor, pixelNumber: int) -> torch.Tensor:
    """(argwhere(lightsource) - pn//2).int()  -- reference imageformation.py:59 (row-major order, values ignored)."""
    return (torch.argwhere(lightsource) - (pixelNumber // 2)).to(torch.int32).contiguous()


def _as_c64(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.complex64, non_blocking=True).contiguous()


class _DevPtr:
    """Minimal __cuda_array_interface__ carrier so that torch can view memory this package allocated itself
    (peer-mapped buffers come from cudaMalloc + CUDA IPC, not from torch's caching allocator)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def tensor_from_ptr(ptr: int, n: int, dev, dtype=torch.float32) -> torch.Tensor:
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevPtr(ptr, n, typestr), device=dev)


def _check_square(name: str, t, pn: int | None = None) -> int:
    """The kernels index every input plane with pitch pn: reject anything that is not [pn, pn] (the reference
    raises a broadcast error for mismatched shapes, imageformation.py:34/:63)."""
    if t.dim() != 2 or t.shape[0] != t.shape[1]:
        raise _native.LithoError(f"{name} must be a square 2-D tensor, got shape {tuple(t.shape)}")
    if pn is not None and t.shape[0] != pn:
        raise _native.LithoError(f"{name} has shape {tuple(t.shape)}, expected ({pn}, {pn}) like maskFT")
    return int(t.shape[0])


class PreparedImage:
    """Inputs of one aerial image staged on the device by AbbeEngine.prepare()."""

    __slots__ = ("slot", "maskFT", "pupil", "shifts", "plan", "eps", "ready")

    def __init__(self, slot, maskFT, pupil, shifts, plan, eps, ready):
        self.slot, self.maskFT, self.pupil, self.shifts = slot, maskFT, pupil, shifts
        self.plan, self.eps, self.ready = plan, eps, ready


class AbbeEngine:
    """Per-device cache of plans and workspaces for the FFT-approximation path.

    Threading: the caches are guarded by a lock, but a plan carries per-call state of its own (T-ring events, an
    auxiliary stream; include/litho_b200.h) and every fast plan owns its T ring, so one engine drives ONE stream of
    images at a time.  Callers that want two concurrent streams on a device use two plans (different pupils get
    different plans anyway) and serialise calls that share a plan."""

    _engines: dict = {}
    _lock = threading.Lock()
    MAX_STAGED_POINTS = 1 << 20     # source points per rank the staging buffers of prepare() hold (8 MB per slot)

    def __init__(self, device: torch.device):
        self.device = device
        self.lib = _native.device_lib()
        self._plans: dict = {}
        self._workspaces: dict = {}
        self._staging: dict = {}        # (slot, pn, source dtype) -> device copies of maskFT, pupil, lightsource
        self._staging_busy: dict = {}   # slot -> event of the last run() that read the set
        self._copy = None
        self._cache_lock = threading.Lock()

    @classmethod
    def get(cls, device) -> "AbbeEngine":
        dev = _require_cuda(device)
        with cls._lock:
            eng = cls._engines.get(dev)
            if eng is None:
                eng = cls._engines[dev] = AbbeEngine(dev)
            return eng

    # -- helpers ---------------------------------------------------------------------------
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def pupil_bbox(self, pupil_d: torch.Tensor):
        return self.lib.pupil_bbox(pupil_d.data_ptr(), pupil_d.shape[0], self.stream())

    def pupil_support(self, pupil_d: torch.Tensor):
        """bbox + rim extents of the non-zero pupil samples (12 ints, one device sync)."""
        return self.lib.pupil_support(pupil_d.data_ptr(), pupil_d.shape[0], self.stream())

    def plan(self, pn: int, N: int, support, generic: bool = False) -> _native.Plan:
        key = (pn, N, tuple(support), bool(generic))
        with self._cache_lock:
            p = self._plans.get(key)
            if p is None:
                p = self._plans[key] = self.lib.plan_create(pn, N, support, _native.PLAN_GENERIC if generic else 0)
        return p

    def plan_for(self, pn: int, N: int, support, shifts_d: torch.Tensor) -> _native.Plan:
        """Fast coarse-grid plan when no source point wraps the pupil window, generic plan otherwise."""
        plan = self.plan(pn, N, support)
        if plan.path == 2:
            n = int(shifts_d.shape[0])
            bounds = self.lib.shift_bounds(shifts_d.data_ptr() if n else None, n, self.stream())
            if not plan.shifts_fit(bounds):
                plan = self.plan(pn, N, support, generic=True)
        return plan

    def workspace(self, nbytes: int, slot: str = "t") -> torch.Tensor:
        with self._cache_lock:
            ws = self._workspaces.get(slot)
            if ws is None or ws.numel() < nbytes:
                ws = self._workspaces[slot] = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        return ws

    # -- hot path --------------------------------------------------------------------------
    def accumulate(self, plan: _native.Plan, maskFT_d, pupil_d, shifts_d, intensity, weights_d=None, batch: int = 0,
                   inputs_ready: bool = False):
        """intensity += sum_s w_s |IDFT{roll(P, shift_s) * M}|^2 (residue-major plane of `plan`).
        inputs_ready: the input tensors are already valid on the device (not produced by work still queued on
        the current stream), so the row pass may overlap the tail of the previous image (LITHO_PHASE_INPUTS_READY)."""
        n_src = int(shifts_d.shape[0])
        if n_src == 0:
            return
        batch = self.batch_for(plan, n_src, batch)
        wsb = plan.workspace_bytes(batch)
        # a fast plan keeps a T ring of its own: its row passes run ahead on the plan's auxiliary stream (chained
        # across images with inputs_ready), so nothing else may ever be queued on that memory
        ws = self.workspace(wsb, ("t", plan.handle.value) if plan.path == 2 else "t")
        pn = plan.pn
        if tuple(maskFT_d.shape) != (pn, pn) or tuple(pupil_d.shape) != (pn, pn):
            raise _native.LithoError(f"accumulate: maskFT {tuple(maskFT_d.shape)} / pupil {tuple(pupil_d.shape)} do not "
                                     f"match the plan's grid ({pn}, {pn})")
        if shifts_d.dim() != 2 or shifts_d.shape[1] != 2 or shifts_d.dtype != torch.int32:
            raise _native.LithoError("accumulate: shifts must be an int32 tensor of shape [n_src, 2]")
        if weights_d is not None and int(weights_d.numel()) != n_src:
            raise _native.LithoError(f"accumulate: {int(weights_d.numel())} weights for {n_src} source points")
        if intensity.numel() < plan.intensity_elems:
            raise _native.LithoError("accumulate: intensity plane smaller than plan.intensity_elems")
        plan.accumulate(maskFT_d.data_ptr(), pupil_d.data_ptr(), shifts_d.data_ptr(),
                        None if weights_d is None else weights_d.data_ptr(), n_src, batch,
                        intensity.data_ptr(), ws.data_ptr(), wsb, self.stream(),
                        phases=3 | (_native.PHASE_INPUTS_READY if inputs_ready else 0))

    @staticmethod
    def batch_for(plan: _native.Plan, n_src: int, batch: int = 0) -> int:
        """Source points per launch pair: the plan's default (sized from measurements), but at least 4 batches per
        call so that row and column passes overlap -- the rule litho_abbe_fft_accumulate applies for batch <= 0."""
        if batch <= 0:
            batch = min(plan.default_batch, max(1, -(-n_src // 4)))
        return max(1, min(batch, n_src))

    def intensity_plane(self, plan: _native.Plan) -> torch.Tensor:
        return torch.zeros(plan.intensity_elems, dtype=torch.float32, device=self.device)

    def finalize(self, plan: _native.Plan, intensity: torch.Tensor, eps: float) -> torch.Tensor:
        side = plan.output_side(eps)
        out = torch.empty((side, side), dtype=torch.float32, device=self.device)
        fwb = plan.finalize_workspace_bytes()
        fws = self.workspace(fwb, "finalize")
        plan.finalize(intensity.data_ptr(), eps, out.data_ptr(), fws.data_ptr(), fwb, self.stream())
        return out

    def unpermute(self, plan: _native.Plan, intensity: torch.Tensor) -> torch.Tensor:
        out = torch.empty((plan.pn, plan.pn), dtype=torch.float32, device=self.device)
        fwb = plan.finalize_workspace_bytes()
        fws = self.workspace(fwb, "finalize")
        plan.unpermute(intensity.data_ptr(), out.data_ptr(), fws.data_ptr(), fwb, self.stream())
        return out

    def abbe_fft(self, maskFT, pupilF, lightsource, pixelSize, deltaK, wavelength, *, weights=None, batch: int = 0,
                 shifts=None, reduce_fn=None, postprocess: bool = True, generic: bool = False,
                 plan: _native.Plan | None = None) -> torch.Tensor:
        """abbeImage(fft=True).  `shifts` (int32 [n,2]) overrides the source-point extraction and
        `reduce_fn(intensity)` runs between accumulation and post-processing (multi-GPU sum).
        `generic` forces the fine-grid kernels; `plan` pins a plan chosen by the caller (all ranks of a
        sharded image must use the same one so that their intensity planes can be summed)."""
        dev = self.device
        with torch.cuda.device(dev):
            pn = _check_square("maskFT", maskFT)
            _check_square("pupilF", pupilF, pn)
            maskFT_d = _as_c64(maskFT, dev)
            pupil_d = _as_c64(pupilF, dev)
            eps, N = epsilon_n(deltaK, pixelSize, wavelength)
            if shifts is None:
                _check_square("lightsource", lightsource, pn)      # shifts are centred with the mask's pn
                ls_d = lightsource.to(dev, non_blocking=True)
                shifts_d = source_shifts(ls_d, pn)
            else:
                shifts_d = shifts.to(device=dev, dtype=torch.int32).contiguous()
                if shifts_d.dim() != 2 or shifts_d.shape[1] != 2:
                    raise _native.LithoError(f"shifts must have shape [n_src, 2], got {tuple(shifts_d.shape)}")
            w_d = None if weights is None else weights.to(device=dev, dtype=torch.float32).contiguous()
            if w_d is not None and int(w_d.numel()) != int(shifts_d.shape[0]):
                raise _native.LithoError(f"{int(w_d.numel())} weights for {int(shifts_d.shape[0])} source points")
            if plan is None:
                support = self.pupil_support(pupil_d)
                plan = self.plan(pn, N, support, generic=True) if generic else self.plan_for(pn, N, support, shifts_d)
            else:
                if plan.pn != pn or plan.N != N:
                    raise _native.LithoError(f"pinned plan is for pn={plan.pn}, N={plan.N}; inputs need pn={pn}, N={N}")
                if plan.path == 2:   # a pinned fast plan must still satisfy the no-wrap contract
                    n = int(shifts_d.shape[0])
                    bounds = self.lib.shift_bounds(shifts_d.data_ptr() if n else None, n, self.stream())
                    if not plan.shifts_fit(bounds):
                        raise _native.LithoError(
                            f"pinned fast plan: source shifts {bounds} leave the plan's no-wrap range "
                            f"{plan.shift_range}; use a generic plan for sources that wrap the pupil window")
            intensity = self.intensity_plane(plan)
            self.accumulate(plan, maskFT_d, pupil_d, shifts_d, intensity, w_d, batch)
            if reduce_fn is not None:
                reduce_fn(intensity)
            return self.finalize(plan, intensity, eps) if postprocess else self.unpermute(plan, intensity)

    def abbe_fft_focus(self, maskFT, pupils, lightsource, pixelSize, deltaK, wavelength, *, weights=None,
                       batch: int = 0, shifts=None, postprocess: bool = True, focus_batch: int = 0):
        """Focus batching (BASELINE cfg5): abbeImage(fft=True) for every pupil function in `pupils` (a list of
        [pn, pn] tensors or one [F, pn, pn] tensor: focus / aberration variants of ONE pupil, i.e. the same support) with
        the mask spectrum and the source shared.  One row pass per batch of source points serves `focus_batch` focus
        values at a time (default: all of them), so each shifted mask-spectrum row is fetched once per group; then one
        column pass per focus value.  Returns the list of images; each equals the single-image call bit for bit."""
        dev = self.device
        with torch.cuda.device(dev):
            pn = _check_square("maskFT", maskFT)
            maskFT_d = _as_c64(maskFT, dev)
            if isinstance(pupils, torch.Tensor):
                if pupils.dim() != 3 or tuple(pupils.shape[1:]) != (pn, pn):
                    raise _native.LithoError(f"pupils must have shape [F, {pn}, {pn}], got {tuple(pupils.shape)}")
                stack = _as_c64(pupils, dev)
            else:
                for q in pupils:
                    _check_square("pupil", q, pn)
                stack = torch.stack([_as_c64(q, dev) for q in pupils])
            F = int(stack.shape[0])
            if F == 0:
                return []
            eps, N = epsilon_n(deltaK, pixelSize, wavelength)
            if shifts is None:
                _check_square("lightsource", lightsource, pn)
                shifts_d = source_shifts(lightsource.to(dev, non_blocking=True), pn)
            else:
                shifts_d = shifts.to(device=dev, dtype=torch.int32).contiguous()
            n_src = int(shifts_d.shape[0])
            w_d = None if weights is None else weights.to(device=dev, dtype=torch.float32).contiguous()
            if w_d is not None and int(w_d.numel()) != n_src:
                raise _native.LithoError(f"{int(w_d.numel())} weights for {n_src} source points")
            # one plan for all focus values: they must share the support (window and rim extents)
            support = self.pupil_support(stack[0])
            for f in range(1, F):
                if self.pupil_support(stack[f]) != support:
                    raise _native.LithoError("abbe_fft_focus: the pupils do not share one support; image them one by one")
            plan = self.plan_for(pn, N, support, shifts_d)
            elems = plan.intensity_elems
            stride = (elems + 63) // 64 * 64
            planes = torch.zeros((F, stride), dtype=torch.float32, device=dev)
            fb = F if focus_batch <= 0 else min(F, focus_batch)
            if n_src:
                for f0 in range(0, F, fb):
                    nf = min(fb, F - f0)
                    wsb = plan.workspace_bytes_focus(batch, nf)
                    ws = self.workspace(wsb, ("t", plan.handle.value) if plan.path == 2 else "t")
                    plan.accumulate_focus(maskFT_d.data_ptr(), stack[f0].data_ptr(), nf, pn * pn, shifts_d.data_ptr(),
                                          None if w_d is None else w_d.data_ptr(), n_src, batch, planes[f0].data_ptr(),
                                          stride, ws.data_ptr(), wsb, self.stream())
            fin = self.finalize if postprocess else (lambda pl, it, e: self.unpermute(pl, it))
            return [fin(plan, planes[f], eps) for f in range(F)]

    # -- pipelined form of abbe_fft: stage inputs one image ahead --------------------------------
    def prepare(self, maskFT, pupilF, lightsource, pixelSize, deltaK, wavelength, *, stream=None, slot: int = 0,
                shard=None, generic: bool = False, plan: _native.Plan | None = None,
                upload_group=None, upload_peers=None) -> "PreparedImage":
        """First half of abbe_fft: upload the (host or device) inputs and extract the source points on `stream`
        (default: a per-engine copy stream), without touching the compute stream.  Host tensors should be
        pinned.  `slot` (0/1) selects one of two device staging sets, so that image i+1 can be prepared while
        image i is being computed; a set is reused only after the run() that consumed it has finished (the
        copy stream waits for that run's event).  `shard = (rank, world)` keeps every world-th source point.
        `upload_group` (a torch.distributed group, all ranks passing the same host inputs): every rank uploads only
        its 1/world slice of each tensor over PCIe and the slices are exchanged with one all-gather per tensor over
        NVLink, instead of every rank pulling the full tensors through the host's memory system at once.  Give it
        a group of its own (dist.new_group) so that it does not queue behind the intensity reduce.
        `upload_peers` (a distributed.PeerStaging sized for the three tensors, all ranks passing the same host inputs):
        the same 1/world upload, but the slices are exchanged by copy-engine transfers between peer-mapped staging
        buffers (no kernel, nothing competes with the compute kernels for SMs); takes precedence over upload_group."""
        dev = self.device
        st = stream if stream is not None else self._copy_stream()
        with torch.cuda.device(dev), torch.cuda.stream(st):
            prev = self._staging_busy.get(slot)
            if prev is not None:
                st.wait_event(prev)
            pn = _check_square("maskFT", maskFT)
            _check_square("pupilF", pupilF, pn)
            _check_square("lightsource", lightsource, pn)
            if plan is not None and plan.pn != pn:
                raise _native.LithoError(f"pinned plan is for pn={plan.pn}, inputs have pn={pn}")
            use_peers = upload_peers is not None and shard is not None and shard[1] > 1
            if use_peers:
                bufs = self._peer_views(upload_peers, slot, pn, lightsource)
            else:
                bufs = self._staging.get((slot, pn, lightsource.dtype))
                if bufs is None:
                    bufs = self._staging[(slot, pn, lightsource.dtype)] = (
                        torch.empty((pn, pn), dtype=torch.complex64, device=dev),
                        torch.empty((pn, pn), dtype=torch.complex64, device=dev),
                        torch.empty((pn, pn), dtype=lightsource.dtype, device=dev))
            mft_d, pf_d, ls_d = bufs[-3:]
            if use_peers:
                self._peer_upload(upload_peers, slot, bufs[0], (maskFT, pupilF, lightsource), st)
            elif upload_group is not None and shard is not None and shard[1] > 1:
                self._sharded_upload(bufs, (maskFT, pupilF, lightsource), shard, upload_group, slot)
            else:
                mft_d.copy_(maskFT, non_blocking=True)
                pf_d.copy_(pupilF, non_blocking=True)
                ls_d.copy_(lightsource, non_blocking=True)
            eps, N = epsilon_n(deltaK, pixelSize, wavelength)
            # source points of this rank's shard + the shift bounds of all points in two short launches (litho_source_points;
            # host sync on the copy stream only).  The torch op sequence costs ~10 small kernels, each of which waits
            # for a free SM slot next to the persistent compute kernels of the image being computed.
            rank_, world_ = shard if shard is not None else (0, 1)
            shifts = bounds = None
            if not ls_d.is_complex():
                sbuf = self._staging.get(("shifts", slot))
                if sbuf is None:
                    sbuf = self._staging[("shifts", slot)] = torch.empty((self.MAX_STAGED_POINTS, 2), dtype=torch.int32, device=dev)
                n_all, n_mine, bounds = self.lib.source_points(ls_d.data_ptr(), ls_d.element_size(), ls_d.is_floating_point(),
                                                               pn, rank_, world_, sbuf.data_ptr(), self.MAX_STAGED_POINTS,
                                                               st.cuda_stream)
                if n_mine <= self.MAX_STAGED_POINTS:
                    shifts = sbuf[:n_mine]
            if shifts is None:      # complex-valued source plane or more points than the staging buffer holds
                shifts_all = source_shifts(ls_d, pn)
                n = int(shifts_all.shape[0])
                bounds = self.lib.shift_bounds(shifts_all.data_ptr() if n else None, n, st.cuda_stream)
                shifts = shifts_all if shard is None else shifts_all[shard[0]::shard[1]].contiguous()
            if plan is None:
                support = self.lib.pupil_support(pf_d.data_ptr(), pn, st.cuda_stream)
                plan = self.plan(pn, N, support, generic=generic)
                if plan.path == 2 and not plan.shifts_fit(bounds):
                    plan = self.plan(pn, N, support, generic=True)
            elif plan.path == 2 and not plan.shifts_fit(bounds):     # pinned fast plan: the no-wrap contract
                raise _native.LithoError(f"pinned fast plan: source shifts {bounds} leave the plan's no-wrap range "
                                         f"{plan.shift_range}")
            ready = torch.cuda.Event()
            ready.record(st)
        return PreparedImage(slot, mft_d, pf_d, shifts, plan, eps, ready)

    def run(self, prep: "PreparedImage", *, batch: int = 0, reduce_fn=None, postprocess: bool = True,
            finalize: bool = True):
        """Second half of abbe_fft on the current stream: accumulate, optional `reduce_fn(intensity)`, post-process.
        With finalize=False the (reduced) intensity plane is left to the caller (ranks that are not the root of a
        rooted reduce) and None is returned."""
        dev = self.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(prep.ready)
            prep.shifts.record_stream(cur)        # allocated on the copy stream, read by kernels of this one
            intensity = self.intensity_plane(prep.plan)
            self.accumulate(prep.plan, prep.maskFT, prep.pupil, prep.shifts, intensity, None, batch)
            self.consumed(prep)                   # the staging set is free once the accumulation has read it
            if reduce_fn is not None:
                reduce_fn(intensity)
            if not finalize:
                return None
            return self.finalize(prep.plan, intensity, prep.eps) if postprocess else self.unpermute(prep.plan, intensity)

    @staticmethod
    def peer_staging_bytes(pn: int, lightsource_dtype=torch.int64) -> int:
        """Bytes a distributed.PeerStaging must hold per slot for prepare(upload_peers=...): maskFT, pupil, lightsource."""
        return 2 * pn * pn * 8 + pn * pn * torch.empty(0, dtype=lightsource_dtype).element_size()

    def _peer_views(self, peers, slot, pn, lightsource):
        """(uint8 view of the slot's buffer, maskFT, pupil, lightsource) as tensors over a PeerStaging buffer."""
        key = ("peerviews", slot, pn, lightsource.dtype, peers.buffer_ptr(slot))
        v = self._staging.get(key)
        if v is None:
            nb = pn * pn * 8
            total = self.peer_staging_bytes(pn, lightsource.dtype)
            if peers.nbytes != total:
                raise _native.LithoError(f"PeerStaging holds {peers.nbytes} bytes per slot, these inputs need {total}")
            u8 = tensor_from_ptr(peers.buffer_ptr(slot), total, self.device, torch.uint8)
            mft_d = torch.view_as_complex(u8[:nb].view(torch.float32).view(pn, pn, 2))
            pf_d = torch.view_as_complex(u8[nb:2 * nb].view(torch.float32).view(pn, pn, 2))
            ls_d = u8[2 * nb:].view(lightsource.dtype).view(pn, pn)
            v = self._staging[key] = (u8, mft_d, pf_d, ls_d)
        return v

    def _peer_upload(self, peers, slot, u8, host_tensors, st):
        """H2D of this rank's byte slice of [maskFT | pupil | lightsource] + copy-engine gather of the other slices."""
        segs, start = [], 0
        for k, src in enumerate(host_tensors):
            src = src.contiguous()
            if k < 2 and src.dtype != torch.complex64:
                src = src.to(torch.complex64)
            flat = torch.view_as_real(src).view(-1).view(torch.uint8) if src.is_complex() else src.view(-1).view(torch.uint8)
            segs.append((start, flat))
            start += flat.numel()

        def upload_my_slice(dst_ptr, off, n):
            for s0, flat in segs:
                a, b = max(off, s0), min(off + n, s0 + flat.numel())
                if a < b:
                    u8[a:b].copy_(flat[a - s0:b - s0], non_blocking=True)

        peers.gather(slot, upload_my_slice, st.cuda_stream)

    def _sharded_upload(self, dev_bufs, host_tensors, shard, group, slot):
        """H2D of this rank's byte slice of each input + all-gather of the slices (current stream = copy stream)."""
        import torch.distributed as dist
        rank, world = shard
        for k, (full, src) in enumerate(zip(dev_bufs, host_tensors)):
            src = src.contiguous()
            if src.dtype != full.dtype:
                src = src.to(full.dtype)
            flat_dst = torch.view_as_real(full).view(-1).view(torch.uint8) if full.is_complex() else full.view(-1).view(torch.uint8)
            flat_src = torch.view_as_real(src).view(-1).view(torch.uint8) if src.is_complex() else src.view(-1).view(torch.uint8)
            n = flat_dst.numel()
            if n % world or flat_src.device.type == "cuda":
                full.copy_(src, non_blocking=True)      # ragged split or already on a device: plain copy
                continue
            chunk = n // world
            key = ("part", slot, k, chunk)
            part = self._staging.get(key)
            if part is None:
                part = self._staging[key] = torch.empty(chunk, dtype=torch.uint8, device=self.device)
            part.copy_(flat_src[rank * chunk:(rank + 1) * chunk], non_blocking=True)
            dist.all_gather_into_tensor(flat_dst, part, group=group)

    def consumed(self, prep: "PreparedImage"):
        """Mark `prep`'s staging set as read by everything enqueued so far on the current stream (for callers
        that drive accumulate() themselves instead of run())."""
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        self._staging_busy[prep.slot] = done

    def _copy_stream(self) -> torch.cuda.Stream:
        if self._copy is None:
            self._copy = torch.cuda.Stream(self.device)
        return self._copy

    def fft_field(self, pf, maskFT, pixelNumber: int, N: int) -> torch.Tensor:
        dev = self.device
        with torch.cuda.device(dev):
            pn = _check_square("maskFFFT", maskFT)
            _check_square("pf", pf, pn)
            pf_d = _as_c64(pf, dev)
            maskFT_d = _as_c64(maskFT, dev)
            plan = self.plan(pn, int(N), self.pupil_bbox(pf_d), generic=True)
            wsb = plan.workspace_bytes(1)
            ws = self.workspace(wsb)
            field = torch.empty((pn, pn), dtype=torch.complex64, device=dev)
            plan.fft_field(pf_d.data_ptr(), maskFT_d.data_ptr(), field.data_ptr(), ws.data_ptr(), wsb, self.stream())
            return field


def calculateFFTAerial(pf: torch.Tensor, maskFFFT: torch.Tensor, pixelNumber: int, N: int) -> torch.Tensor:
    """Gau-2023 FFT-approximation field of one (already shifted) pupil -- reference imageformation.py:32-45."""
    dev = maskFFFT.device if maskFFFT.is_cuda else (pf.device if pf.is_cuda else None)
    return AbbeEngine.get(dev).fft_field(pf, maskFFFT, pixelNumber, N)


def calculateAerial(pupil: torch.Tensor, maskFT: torch.Tensor, fraunhoferConstant, pixelNumber: int, pixelSize,
                    device) -> torch.Tensor:
    """Direct ("Abbe") solver field -- reference imageformation.py:3-30."""
    from .direct import direct_field  # local import: separate kernel family
    return direct_field(pupil, maskFT, fraunhoferConstant, pixelNumber, pixelSize, _require_cuda(device))


def abbeImage(mask, maskFT: torch.Tensor, pupilF: torch.Tensor, lightsource: torch.Tensor, pixelSize, deltaK: float,
              wavelength: float, fft: bool, device) -> torch.Tensor:
    """Partially coherent aerial image by Abbe source-point summation -- reference imageformation.py:47-77.

    Same arguments and result as the reference (float32 image on `device`); `mask` is accepted
    for signature compatibility (the reference only uses it to reach calculateEpsilonN).
    """
    dev = _require_cuda(device)
    if fft:
        return AbbeEngine.get(dev).abbe_fft(maskFT, pupilF, lightsource, pixelSize, deltaK, wavelength)
    from .direct import direct_abbe_image
    return direct_abbe_image(maskFT, pupilF, lightsource, pixelSize, wavelength, dev)
