"""Multi-GPU Abbe imaging: source points sharded across ranks, one sum-reduce of the intensity plane.

I = sum_s |E_s|^2 has independent terms (reference imageformation.py:62-67), so the source-point list
is split across the ranks of one torch.distributed process group (one process per GPU, NCCL over
NVLink/NVSwitch), every rank accumulates a partial intensity plane with the SAME plan, the planes are
summed with a single all-reduce and the post-processing runs after the reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .imaging import AbbeEngine, _as_c64, _require_cuda, epsilon_n, source_shifts

__all__ = ["shard_shifts", "abbe_image_sharded", "focus_sweep_sharded"]


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
                        batch: int = 0):
    """Focus-exposure sweep (BASELINE cfg5): one aerial image per pupil function in `pupils`, the pupils
    (focus values) sharded across the ranks -- independent images, so no reduce; every rank returns the
    full list (images of other ranks are received with one all_gather per image slot).

    With a single process it is simply a loop over the pupils that reuses the uploaded mask spectrum and
    source-point list."""
    dev = _require_cuda(device)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    eng = AbbeEngine.get(dev)
    with torch.cuda.device(dev):
        maskFT_d = _as_c64(maskFT, dev)
        pn = int(maskFT_d.shape[0])
        eps, N = epsilon_n(deltaK, pixelSize, wavelength)
        shifts_d = source_shifts(lightsource.to(dev), pn)
        mine = {}
        for i in range(rank, len(pupils), world):
            pupil_d = _as_c64(pupils[i], dev)
            plan = eng.plan_for(pn, N, eng.pupil_support(pupil_d), shifts_d)
            intensity = eng.intensity_plane(plan)
            eng.accumulate(plan, maskFT_d, pupil_d, shifts_d, intensity, None, batch)
            mine[i] = eng.finalize(plan, intensity, eps)
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
