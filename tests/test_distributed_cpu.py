"""World-size-2 check of the multi-GPU scheme on CPU (gloo): source points sharded across ranks, each rank
accumulates a partial intensity plane through the C ABI (CPU emulation of the kernels -- test
infrastructure), one sum all-reduce, post-processing after the reduce.  Must equal the single-process image
and the reference golden."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from oracle import abbe_oracle as O
from lithographysimulator_b200.distributed import shard_shifts


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = H.load_kat()[case]
        lib = H.emu_lib()
        pn = c["maskFT"].shape[0]
        ps = float(c["pixel_size"])
        eps, N = lib.epsilon_n(4 / pn, ps, 193.0)
        mft = np.ascontiguousarray(c["maskFT"])
        pup = np.ascontiguousarray(c["pupil"])
        shifts_all = np.ascontiguousarray(O.source_shifts(c["lightsource"], pn))
        support = lib.pupil_support(pup.ctypes.data, pn)
        # plan chosen from ALL source points so that both ranks use the same intensity layout
        plan = lib.plan_create(pn, N, support)
        if plan.path == 2 and not plan.shifts_fit(lib.shift_bounds(shifts_all.ctypes.data, len(shifts_all))):
            plan = lib.plan_create(pn, N, support, 1)
        mine = np.ascontiguousarray(shard_shifts(torch.from_numpy(shifts_all), rank, world).numpy())
        inten = np.zeros(plan.intensity_elems, np.float32)
        wsb = plan.workspace_bytes(0)
        ws = np.zeros(max(wsb, 8), np.uint8)
        plan.accumulate(mft.ctypes.data, pup.ctypes.data, mine.ctypes.data, None, len(mine), 0, inten.ctypes.data,
                        ws.ctypes.data, wsb)
        t = torch.from_numpy(inten)
        dist.all_reduce(t)  # sum of the partial planes (NCCL over NVLink on the GPU box)
        side = plan.output_side(eps)
        out = np.zeros((side, side), np.float32)
        fwb = plan.finalize_workspace_bytes()
        fws = np.zeros(max(fwb, 8), np.uint8)
        plan.finalize(inten.ctypes.data, eps, out.ctypes.data, fws.ctypes.data, fwb)
        if rank == 0:
            np.save(out_path, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["demo64_quasar", "wrap_128"])
def test_two_rank_sharded_image_matches_reference(tmp_path, case):
    H.emu_lib()  # build once before forking
    out_path = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, _free_port(), case, out_path), nprocs=2, join=True)
    img = np.load(out_path)
    ref = H.load_kat()[case]["image"]
    assert img.shape == ref.shape
    assert O.rel_l2(img, ref) < H.TOL


def test_shard_shifts_partition():
    sh = torch.arange(46).reshape(23, 2)
    parts = [shard_shifts(sh, r, 4) for r in range(4)]
    assert sum(len(p) for p in parts) == 23 and max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    merged = torch.cat(parts).tolist()
    assert sorted(merged) == sh.tolist()


def _peer_worker(rank, world, port, case, out_dir):
    """The peer-memory protocol of distributed.PeerPlanes (publish / gather_sum / wait_consumed) between two
    PROCESSES, on the CPU emulation (POSIX shared memory stands in for CUDA IPC): 7 images over 3 plane slots with a rotating root,
    every image must equal the reference golden."""
    import ctypes as C
    from lithographysimulator_b200.distributed import PeerPlanes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = H.load_kat()[case]
        lib = H.emu_lib()
        pn = c["maskFT"].shape[0]
        eps, N = lib.epsilon_n(4 / pn, float(c["pixel_size"]), 193.0)
        mft = np.ascontiguousarray(c["maskFT"])
        pup = np.ascontiguousarray(c["pupil"])
        shifts_all = np.ascontiguousarray(O.source_shifts(c["lightsource"], pn))
        plan = lib.plan_create(pn, N, lib.pupil_support(pup.ctypes.data, pn))
        if plan.path == 2 and not plan.shifts_fit(lib.shift_bounds(shifts_all.ctypes.data, len(shifts_all))):
            plan = lib.plan_create(pn, N, lib.pupil_support(pup.ctypes.data, pn), 1)
        mine = np.ascontiguousarray(shard_shifts(torch.from_numpy(shifts_all), rank, world).numpy())
        elems = plan.intensity_elems

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        SLOTS = 3                                  # what distributed.ShardedPipeline uses
        peers = PeerPlanes(lib, elems, rank, world, exchange, slots=SLOTS)
        planes = [np.ctypeslib.as_array((C.c_float * elems).from_address(peers.plane_ptr(k))) for k in range(SLOTS)]
        wsb = plan.workspace_bytes(0)
        ws = np.zeros(max(wsb, 8), np.uint8)
        fwb = plan.finalize_workspace_bytes()
        fws = np.zeros(max(fwb, 8), np.uint8)
        side = plan.output_side(eps)
        summed = np.zeros(elems, np.float32)
        for i in range(7):
            k, seq, root = i % SLOTS, i + 1, i % world
            if i >= SLOTS:
                peers.wait_consumed(k, seq - SLOTS)
            planes[k][:] = 0
            plan.accumulate(mft.ctypes.data, pup.ctypes.data, mine.ctypes.data, None, len(mine), 0,
                            peers.plane_ptr(k), ws.ctypes.data, wsb)
            peers.publish(k, seq, root)
            if rank == root:
                peers.gather_sum(k, seq, summed.ctypes.data)
                out = np.zeros((side, side), np.float32)
                plan.finalize(summed.ctypes.data, eps, out.ctypes.data, fws.ctypes.data, fwb)
                np.save(os.path.join(out_dir, f"img{i}.npy"), out)
        err = (C.c_int * 2).from_address(peers.err_ptr)
        assert err[0] == 0
        dist.barrier()
        peers.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_peer_memory_sum_rotating_root(tmp_path):
    H.emu_lib()
    mp.spawn(_peer_worker, args=(2, _free_port(), "demo64_quasar", str(tmp_path)), nprocs=2, join=True)
    ref = H.load_kat()["demo64_quasar"]["image"]
    for i in range(7):
        img = np.load(str(tmp_path / f"img{i}.npy"))
        assert O.rel_l2(img, ref) < H.TOL


def _staging_worker(rank, world, port, out_dir):
    """distributed.PeerStaging between processes on the CPU emulation: every rank contributes its slice of a byte
    range whose length is not a multiple of anything, 5 uses over 2 slots; every rank must end up with all bytes."""
    import ctypes as C
    from lithographysimulator_b200.distributed import PeerStaging
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = H.emu_lib()
        nbytes = 100_003

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        st = PeerStaging(lib, nbytes, rank, world, exchange)
        ok = True
        for use in range(5):
            slot = use % 2
            data = ((np.arange(nbytes, dtype=np.int64) * (use + 3) + 7 * use) % 251).astype(np.uint8)   # same on all ranks

            def upload(dst_ptr, off, n, data=data):
                C.memmove(dst_ptr, data[off:off + n].ctypes.data, n)

            st.gather(slot, upload)
            got = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(st.buffer_ptr(slot)))
            ok = ok and bool((got == data).all())
        err = (C.c_int * 2).from_address(st.err_ptr)
        ok = ok and err[0] == 0
        dist.barrier()
        st.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(ok))
    finally:
        dist.destroy_process_group()


def test_peer_staging_gathers_all_slices(tmp_path):
    H.emu_lib()
    mp.spawn(_staging_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    for r in range(3):
        assert open(str(tmp_path / f"ok{r}")).read() == "True"


def _unavailable_worker(rank, world, port, out_dir):
    """A rank that cannot allocate (or map) its peer buffer: EVERY rank must get PeerUnavailable, after the same number
    of collective calls, so that callers can fall back to the NCCL path together."""
    from lithographysimulator_b200.distributed import PeerPlanes, PeerUnavailable
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = H.emu_lib()

        def exchange(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        class Broken:                       # rank 1's allocation fails, rank 0's mapping of it therefore never happens
            def __getattr__(self, name):
                return getattr(lib, name)

            def peer_alloc(self, nbytes):
                raise RuntimeError("no IPC here")

        got = "none"
        try:
            PeerPlanes(Broken() if rank == 1 else lib, 1000, rank, world, exchange)
        except PeerUnavailable as e:
            got = "unavailable:" + str(e)
        dist.barrier()                      # still in step with each other
        ok = PeerPlanes(lib, 1000, rank, world, exchange)      # and a healthy set-up works right after
        ok.close()
        open(os.path.join(out_dir, f"r{rank}"), "w").write(got)
    finally:
        dist.destroy_process_group()


def test_peer_unavailable_is_raised_on_every_rank(tmp_path):
    H.emu_lib()
    mp.spawn(_unavailable_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = open(str(tmp_path / f"r{r}")).read()
        assert got.startswith("unavailable:") and "no IPC here" in got, got
