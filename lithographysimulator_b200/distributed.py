"""Multi-GPU Abbe imaging: source points sharded across ranks, one sum-reduce of the intensity plane.

I = sum_s |E_s|^2 has independent terms (reference imageformation.py:62-67), so the source-point list
is split across the ranks of one torch.distributed process group (one process per GPU, NCCL over
NVLink/NVSwitch), every rank accumulates a partial intensity plane with the SAME plan, the planes are
summed with a single all-reduce and the post-processing runs after the reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .imaging import AbbeEngine, _as_c64, _require_cuda, epsilon_n, source_shifts, tensor_from_ptr

__all__ = ["shard_shifts", "abbe_image_sharded", "focus_sweep_sharded", "PeerPlanes", "PeerStaging", "PeerUnavailable",
           "ShardedPipeline", "tensor_from_ptr"]


class PeerUnavailable(RuntimeError):
    """Peer-mapped buffers could not be set up on every rank (raised on ALL ranks, so callers can fall back together)."""


def _map_peers(lib, nbytes: int, rank: int, world: int, exchange):
    """Allocate a peer-mappable buffer of nbytes on every rank and map everybody else's.  Collective (three calls of
    `exchange`); a failure on any rank raises PeerUnavailable on every rank after cleaning up."""
    base, handle, err = 0, None, None
    try:
        base, handle = lib.peer_alloc(nbytes)
    except Exception as e:          # noqa: BLE001 -- reported to the other ranks below
        err = f"rank {rank}: {e}"
    handles = exchange(handle)
    bases = [0] * world
    if err is None and all(h is not None for h in handles):
        try:
            for r in range(world):
                bases[r] = base if r == rank else lib.peer_open(handles[r])
        except Exception as e:      # noqa: BLE001
            err = f"rank {rank}: {e}"
    elif err is None:
        err = "a peer could not allocate"
    errs = exchange(err)
    if any(e is not None for e in errs):
        for r, b in enumerate(bases):
            if r != rank and b:
                try:
                    lib.check_peer(lib.litho_peer_close(b), "litho_peer_close")
                except Exception:   # noqa: BLE001
                    pass
        exchange(b"unmapped")       # nobody frees while somebody may still be unmapping
        if base:
            try:
                lib.check_peer(lib.litho_peer_free(base), "litho_peer_free")
            except Exception:       # noqa: BLE001
                pass
        raise PeerUnavailable("; ".join(e for e in errs if e is not None))
    exchange(b"mapped")             # nobody proceeds (or frees) before every rank has mapped every buffer
    return base, bases


class PeerPlanes:
    """`slots` partial-intensity planes per rank in memory every rank of the box maps (CUDA IPC), plus a mailbox of
    64-bit sequence flags, for the peer-memory sum of include/litho_b200.h (litho_peer_*).

    Protocol for image number `seq` (1, 2, ...) in slot k whose root is rank R:
      every rank   accumulates into plane k, then publish(k, seq, R)          -> arrive[k][rank] on R's mailbox
      rank R       gather_sum(k, seq, out): waits for all arrive flags, sums the planes of all ranks in rank order
                   over NVLink into `out`, then writes consumed[k] = seq on every rank's mailbox
      every rank   wait_consumed(k, seq) before it zeroes / reuses plane k
    `exchange(obj)` must return the list of every rank's obj (e.g. a torch.distributed.all_gather_object wrapper);
    the constructor is collective.  Pointers are plain ints, so the same class drives the CPU emulation in the
    world-size-2 gloo tests."""

    MAILBOX_BYTES = 4096

    def __init__(self, lib, elems: int, rank: int, world: int, exchange, slots: int = 2):
        from ._native import MAX_PEERS
        if world > MAX_PEERS:
            raise ValueError(f"PeerPlanes supports at most {MAX_PEERS} ranks")
        self.lib, self.elems, self.rank, self.world, self.slots = lib, int(elems), rank, world, slots
        self.max_peers = MAX_PEERS
        self.plane_bytes = (self.elems * 4 + 255) // 256 * 256
        self.total_bytes = slots * self.plane_bytes + self.MAILBOX_BYTES
        self.base, self.bases = _map_peers(lib, self.total_bytes, rank, world, exchange)

    def plane_ptr(self, slot: int, r: int | None = None) -> int:
        return self.bases[self.rank if r is None else r] + slot * self.plane_bytes

    def _mail(self, r: int) -> int:
        return self.bases[r] + self.slots * self.plane_bytes

    def _arrive_ptr(self, owner: int, slot: int, src: int) -> int:
        return self._mail(owner) + 8 * (slot * self.max_peers + src)

    def _consumed_ptr(self, owner: int, slot: int) -> int:
        return self._mail(owner) + 8 * (self.slots * self.max_peers + slot)

    @property
    def err_ptr(self) -> int:
        return self._mail(self.rank) + self.MAILBOX_BYTES - 8

    def publish(self, slot: int, seq: int, root: int, stream: int = 0):
        self.lib.peer_signal([self._arrive_ptr(root, slot, self.rank)], seq, stream)

    def gather_sum(self, slot: int, seq: int, out_ptr: int, stream: int = 0):
        flags = self._arrive_ptr(self.rank, slot, 0)
        self.lib.peer_wait(flags, self.world, seq, self.err_ptr, stream)
        self.lib.peer_sum(out_ptr, [self.plane_ptr(slot, r) for r in range(self.world)], self.elems, flags, seq,
                          self.err_ptr, stream)
        self.lib.peer_signal([self._consumed_ptr(r, slot) for r in range(self.world)], seq, stream)

    def wait_consumed(self, slot: int, seq: int, stream: int = 0):
        self.lib.peer_wait(self._consumed_ptr(self.rank, slot), 1, seq, self.err_ptr, stream)

    def close(self):
        for r, b in enumerate(self.bases):
            if r != self.rank and b:
                self.lib.check_peer(self.lib.litho_peer_close(b), "litho_peer_close")
        if self.base:
            self.lib.check_peer(self.lib.litho_peer_free(self.base), "litho_peer_free")
        self.base, self.bases = 0, []


class PeerStaging:
    """Input staging when all N ranks of a box need the same host inputs (one image computed N-way): every rank
    uploads 1/N of the bytes over its own PCIe link into its peer-mapped buffer and pulls the other slices from the
    peers' buffers with copy-engine transfers over NVLink (litho_peer_copy: no kernel, so nothing competes with the
    persistent compute kernels for SMs -- an NCCL all-gather of the slices does).  `slots` buffers of `nbytes` each;
    per slot and use number `seq` = 1, 2, ...:
        wait pulled[slot][*] >= seq-1 (every peer has copied my previous slice)  ->  H2D of my slice
        -> ready[slot][me] = seq on every rank  ->  wait ready[slot][*] >= seq  ->  copy the other slices
        -> pulled[slot][me] = seq on every rank.
    Pointers are plain ints (the CPU emulation drives the same class in the gloo tests)."""

    MAILBOX_BYTES = 4096

    def __init__(self, lib, nbytes: int, rank: int, world: int, exchange, slots: int = 2):
        from ._native import MAX_PEERS
        if world > MAX_PEERS:
            raise ValueError(f"PeerStaging supports at most {MAX_PEERS} ranks")
        self.lib, self.rank, self.world, self.slots = lib, rank, world, slots
        self.max_peers = MAX_PEERS
        self.nbytes = int(nbytes)
        self.chunk = (-(-self.nbytes // world) + 255) // 256 * 256          # slice size, 256-byte aligned
        self.slot_bytes = self.chunk * world
        self.total_bytes = slots * self.slot_bytes + self.MAILBOX_BYTES
        self.base, self.bases = _map_peers(lib, self.total_bytes, rank, world, exchange)
        self.uses = [0] * slots

    def buffer_ptr(self, slot: int, r: int | None = None) -> int:
        return self.bases[self.rank if r is None else r] + slot * self.slot_bytes

    def _mail(self, r: int) -> int:
        return self.bases[r] + self.slots * self.slot_bytes

    def _ready_ptr(self, owner: int, slot: int, src: int) -> int:
        return self._mail(owner) + 8 * (slot * self.max_peers + src)

    def _pulled_ptr(self, owner: int, slot: int, src: int) -> int:
        return self._mail(owner) + 8 * ((self.slots + slot) * self.max_peers + src)

    @property
    def err_ptr(self) -> int:
        return self._mail(self.rank) + self.MAILBOX_BYTES - 8

    def my_slice(self):
        """(offset, length) of this rank's slice of the logical byte range [0, nbytes)."""
        lo = self.rank * self.chunk
        return lo, max(0, min(self.nbytes, lo + self.chunk) - lo)

    def gather(self, slot: int, upload_my_slice, stream: int = 0):
        """`upload_my_slice(dst_ptr, offset, length)` must queue the host-to-device copy of bytes [offset, offset+length)
        of the inputs to dst_ptr on `stream`.  Afterwards (in stream order) the slot's buffer holds all nbytes."""
        self.uses[slot] += 1
        seq = self.uses[slot]
        lib, me, W = self.lib, self.rank, self.world
        lib.peer_wait(self._pulled_ptr(me, slot, 0), W, seq - 1, self.err_ptr, stream)
        lo, n = self.my_slice()
        if n:
            upload_my_slice(self.buffer_ptr(slot) + lo, lo, n)
        lib.peer_signal([self._ready_ptr(r, slot, me) for r in range(W)], seq, stream)
        lib.peer_wait(self._ready_ptr(me, slot, 0), W, seq, self.err_ptr, stream)
        for d in range(1, W):                       # staggered: rank r starts with r+1, so no peer is hit by all at once
            r = (me + d) % W
            off = r * self.chunk
            m = max(0, min(self.nbytes, off + self.chunk) - off)
            if m:
                lib.peer_copy(self.buffer_ptr(slot) + off, self.buffer_ptr(slot, r) + off, m, stream)
        lib.peer_signal([self._pulled_ptr(r, slot, me) for r in range(W)], seq, stream)

    def close(self):
        for r, b in enumerate(self.bases):
            if r != self.rank and b:
                self.lib.check_peer(self.lib.litho_peer_close(b), "litho_peer_close")
        if self.base:
            self.lib.check_peer(self.lib.litho_peer_free(self.base), "litho_peer_free")
        self.base, self.bases = 0, []


class ShardedPipeline:
    """Throughput mode of the sharded imager: every image is accumulated by all ranks
    together (source points interleaved), the partial planes are summed by rank `i mod world` -- which alone
    post-processes image i, on a second stream -- while all ranks already accumulate image i+1.

    reduce = "peer": the root reads the other ranks' planes over NVLink with litho_peer_sum (no collective kernel
    to co-schedule; deterministic rank-order sum); "nccl": one asynchronous ncclReduce per image (the baseline).
    Everything a step needs is allocated here, none of it inside submit()."""

    def __init__(self, eng: AbbeEngine, plan, eps: float, *, group=None, reduce: str = "peer", slots: int | None = None,
                 fin_priority: int | None = None):
        import os
        self.eng, self.plan, self.eps, self.group = eng, plan, eps, group
        # plane slots: an image's plane is reused `slots` images later, after its root has read it
        self.slots = int(os.environ.get("LITHO_PLANE_SLOTS", "3")) if slots is None else slots
        # the summing / post-processing stream outranks the accumulation: its few CTAs take the next free SM slots
        # instead of queueing behind the second wave of a column pass, so planes are released early
        prio = int(os.environ.get("LITHO_FIN_PRIORITY", "-1")) if fin_priority is None else fin_priority
        dev = self.dev = eng.device
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.reduce = reduce if self.world > 1 else "none"
        self.i = 0
        self.fin_stream = torch.cuda.Stream(dev, priority=prio)
        elems = plan.intensity_elems
        self.peers = None
        if self.reduce == "peer":
            def exchange(obj):
                out = [None] * self.world
                dist.all_gather_object(out, obj, group=group)
                return out
            try:
                self.peers = PeerPlanes(eng.lib, elems, self.rank, self.world, exchange, slots=self.slots)
            except PeerUnavailable as e:      # raised on every rank: all fall back to the collective together
                if self.rank == 0:
                    import sys
                    print(f"ShardedPipeline: peer-mapped planes unavailable ({e}); using ncclReduce", file=sys.stderr, flush=True)
                self.reduce = "nccl"
        if self.reduce == "peer":
            torch.cuda.synchronize(dev)
            self.planes = [tensor_from_ptr(self.peers.plane_ptr(k), elems, dev) for k in range(self.slots)]
            self.err = tensor_from_ptr(self.peers.err_ptr, 2, dev, torch.int32)
            self.summed = torch.zeros(elems, dtype=torch.float32, device=dev)
        else:
            self.planes = [eng.intensity_plane(plan) for _ in range(self.slots)]
            self.summed = None
        side = plan.output_side(eps)
        self.images = [torch.zeros((side, side), dtype=torch.float32, device=dev) for _ in range(2)]
        self.fwb = plan.finalize_workspace_bytes()
        self.fws = torch.empty(max(self.fwb, 16), dtype=torch.uint8, device=dev)
        self.fin_done = [None] * self.slots
        self.reduce_work = [None] * self.slots
        self.last_image = None
        self.trace = None    # set to [] to record CUDA events per image (see trace_ms)
        if self.reduce == "nccl":
            # NCCL sets up its channels per (collective, root) on first use: touch every root once
            for r in range(self.world):
                dist.reduce(self.planes[0], dst=r, group=group)
            self.planes[0].zero_()
        # lazy kernel loading of the post-processing path, on its own stream, on EVERY rank
        with torch.cuda.stream(self.fin_stream):
            self._finalize(self.planes[0], self.images[0])
        torch.cuda.synchronize(dev)

    def _finalize(self, plane: torch.Tensor, out: torch.Tensor):
        self.plan.finalize(plane.data_ptr(), self.eps, out.data_ptr(), self.fws.data_ptr(), self.fwb,
                           torch.cuda.current_stream(self.dev).cuda_stream)

    def submit(self, maskFT_d, pupil_d, shifts_mine, *, weights_d=None, batch: int = 0, inputs_ready: bool = False,
               wait_event=None, on_accumulated=None, out_host=None):
        """Queue image number self.i.  `shifts_mine` is this rank's shard of the source points.  wait_event: an event
        the accumulation must wait for (inputs staged on another stream); on_accumulated(): called right after the
        accumulation has been queued (to release staging buffers); out_host: pinned tensor the root copies the image to."""
        i, k = self.i, self.i % self.slots
        self.i += 1
        seq = i + 1
        dev, eng = self.dev, self.eng
        main = torch.cuda.current_stream(dev)
        inten = self.planes[k]
        tr = None
        if self.trace is not None:
            tr = {"i": i, "root": i % self.world}
            self.trace.append(tr)
            tr["begin"] = self._mark(main)
        if self.reduce == "peer":
            if i >= self.slots:   # the root of the image that last used this plane has read it
                self.peers.wait_consumed(k, seq - self.slots, main.cuda_stream)
        else:
            if self.reduce_work[k] is not None:
                self.reduce_work[k].wait()
            if self.fin_done[k] is not None:
                main.wait_event(self.fin_done[k])
        inten.zero_()
        if wait_event is not None:
            main.wait_event(wait_event)
        eng.accumulate(self.plan, maskFT_d, pupil_d, shifts_mine, inten, weights_d, batch, inputs_ready=inputs_ready)
        if on_accumulated is not None:
            on_accumulated()
        root = i % self.world
        work = None
        if self.reduce == "peer":
            self.peers.publish(k, seq, root, main.cuda_stream)
        elif self.reduce == "nccl":
            work = dist.reduce(inten, dst=root, group=self.group, async_op=True)
            self.reduce_work[k] = work
        if tr is not None:
            tr["accumulated"] = self._mark(main)
        if self.rank != root:
            return
        ready = torch.cuda.Event()
        ready.record(main)
        fin = self.fin_stream
        with torch.cuda.stream(fin):
            fin.wait_event(ready)            # never spin on this GPU's own accumulation: order it with an event
            src = inten
            if tr is not None:
                tr["fin_begin"] = self._mark(fin)
            if self.reduce == "peer":
                self.peers.gather_sum(k, seq, self.summed.data_ptr(), fin.cuda_stream)
                src = self.summed
            elif work is not None:
                work.wait()
            if tr is not None:
                tr["summed"] = self._mark(fin)
            img = self.images[(i // self.world) % 2]     # this rank's turns as root alternate between two buffers
            self._finalize(src, img)
            if tr is not None:
                tr["finalized"] = self._mark(fin)
            if out_host is not None:
                out_host.copy_(img, non_blocking=True)
            done = torch.cuda.Event()
            done.record(fin)
            self.fin_done[k] = done
        self.last_image = img

    @staticmethod
    def _mark(stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def trace_ms(self):
        """Recorded events as milliseconds since the first image's begin (call after a synchronize)."""
        if not self.trace:
            return []
        base = self.trace[0]["begin"]
        return [{k: (round(base.elapsed_time(v), 4) if isinstance(v, torch.cuda.Event) else v) for k, v in t.items()}
                for t in self.trace]

    def join(self):
        for w in self.reduce_work:
            if w is not None:
                w.wait()
        torch.cuda.current_stream(self.dev).wait_stream(self.fin_stream)

    def check(self):
        """Raise if a peer wait timed out (call after a synchronize)."""
        if self.peers is not None and int(self.err[0].item()) != 0:
            raise RuntimeError(f"peer-memory sum: wait for a peer timed out (error word {int(self.err[0].item())})")

    def close(self):
        if self.peers is not None:
            torch.cuda.synchronize(self.dev)
            if self.world > 1:
                dist.barrier(group=self.group)
            self.planes, self.err = [], None
            self.peers.close()
            self.peers = None


def shard_shifts(shifts: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Interleaved shard of the [n,2] shift list: rank r takes points r, r+world, ...  Interleaving keeps
    the per-rank work equal to within one source point for any source shape."""
    return shifts[rank::world].contiguous()


def abbe_image_sharded(mask, maskFT, pupilF, lightsource, pixelSize, deltaK, wavelength, device, *, group=None,
                       weights=None, batch: int = 0, postprocess: bool = True) -> torch.Tensor:
    """abbeImage(fft=True) computed by all ranks of `group` together; every rank returns the full image.

    All ranks must pass the same inputs (they are replicated: 2 x 8*pn^2 bytes).  Works unchanged with a
    single process (no process group initialised)."""
    dev = _require_cuda(device)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    eng = AbbeEngine.get(dev)
    with torch.cuda.device(dev):
        maskFT_d = _as_c64(maskFT, dev)
        pupil_d = _as_c64(pupilF, dev)
        pn = int(maskFT_d.shape[0])
        eps, N = epsilon_n(deltaK, pixelSize, wavelength)
        shifts_all = source_shifts(lightsource.to(dev), pn)
        # the plan (fast vs generic kernels, hence the layout of the intensity plane) is chosen from
        # ALL source points so that every rank makes the same choice
        plan = eng.plan_for(pn, N, eng.pupil_support(pupil_d), shifts_all)
        mine = shard_shifts(shifts_all, rank, world)
        w_d = None
        if weights is not None:
            w_d = weights.to(device=dev, dtype=torch.float32)[rank::world].contiguous()
        intensity = eng.intensity_plane(plan)
        eng.accumulate(plan, maskFT_d, pupil_d, mine, intensity, w_d, batch)
        if world > 1:
            dist.all_reduce(intensity, op=dist.ReduceOp.SUM, group=group)
        return eng.finalize(plan, intensity, eps) if postprocess else eng.unpermute(plan, intensity)


def focus_sweep_sharded(mask, maskFT, pupils, lightsource, pixelSize, deltaK, wavelength, device, *, group=None,
                        batch: int = 0, focus_batch: int = 0):
    """Focus-exposure sweep (BASELINE cfg5): one aerial image per pupil function in `pupils`, the pupils
    (focus values) sharded across the ranks -- independent images, so no reduce; every rank returns the
    full list (images of other ranks are received with one all_gather per image slot).

    The focus values of one rank are batched (AbbeEngine.abbe_fft_focus; `focus_batch` caps how many share one row
    pass, 0 = all of them)."""
    dev = _require_cuda(device)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    eng = AbbeEngine.get(dev)
    with torch.cuda.device(dev):
        maskFT_d = _as_c64(maskFT, dev)
        pn = int(maskFT_d.shape[0])
        eps, N = epsilon_n(deltaK, pixelSize, wavelength)
        shifts_d = source_shifts(lightsource.to(dev), pn)
        # this rank's focus values are imaged together: one row pass per batch of source points serves all of them
        # (litho_abbe_fft_accumulate_focus), so the shifted mask-spectrum window is fetched once per group
        idx = list(range(rank, len(pupils), world))
        mine = {}
        if idx:
            imgs = eng.abbe_fft_focus(maskFT_d, [pupils[i] for i in idx], None, pixelSize, deltaK, wavelength,
                                      shifts=shifts_d, batch=batch, focus_batch=focus_batch)
            mine = dict(zip(idx, imgs))
        if world == 1:
            return [mine[i] for i in range(len(pupils))]
        side = eng.plan(pn, N, (0, pn - 1, 0, pn - 1), generic=True).output_side(eps)
        out = []
        rounds = (len(pupils) + world - 1) // world
        for rnd in range(rounds):
            i = rnd * world + rank
            local = mine.get(i)
            if local is None:
                local = torch.zeros((side, side), dtype=torch.float32, device=dev)
            bufs = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(bufs, local, group=group)
            for r in range(world):
                if rnd * world + r < len(pupils):
                    out.append(bufs[r])
        return out
