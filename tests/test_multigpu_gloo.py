"""N > 1 host logic on CPU: world_size-2 and -3 gloo groups shard a frame by interleaved tile rows,
render their shards (oracle stands in for the device here), combine on rank 0 and reproduce the
single-rank frame bit for bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from houdini_gsplat_renderer_b200 import multigpu as M, scene as S
    from oracle import oracle as O
    cl = S.make_cloud(4000, 77, sh=True, scale_mult=2.0)
    fr = S.orbit_frame(168, 100, 40.0)               # ragged: 7 tile rows, last one 4 px tall
    cam = O.camera_from_view(fr.view)
    group = 2 if world == 2 else 1
    mine = O.pipeline(O.make_frame(fr, cam, cl.barycentre(), 3, row_rank=rank, row_world=world, row_group=group), cl)
    rows = M.owned_scanlines(fr.height, rank, world, group)
    assert not mine["rgba"][~rows].any()             # un-owned scanlines stay zero
    fb = torch.from_numpy(mine["rgba"].copy())
    M.combine_on_root(fb, rank, world)
    if rank == 0:
        full = O.pipeline(O.make_frame(fr, cam, cl.barycentre(), 3), cl)
        q.put(bool(np.array_equal(fb.numpy(), full["rgba"])) and mine["n_visible"] < full["n_visible"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_row_shards_combine_exactly(world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_row_ownership_partitions_the_frame():
    from houdini_gsplat_renderer_b200 import multigpu as M
    for h, world, group in [(1080, 8, 1), (100, 3, 2), (16, 4, 1), (4320, 8, 4), (2160, 2, 8)]:
        cover = np.zeros(h, int)
        rows = []
        for r in range(world):
            cover += M.owned_scanlines(h, r, world, group)
            rows += M.owned_tile_rows(h, r, world, group)
        assert np.all(cover == 1) and sorted(rows) == list(range((h + 15) // 16))
    assert M.default_row_group(4320, 8) == 4 and M.default_row_group(1080, 8) == 1 and M.default_row_group(1080, 2) == 4


def _shm_worker(rank, world, port, q):
    """The host-direct hand-off of multigpu.RowPartitionedRenderer with the device taken out: every rank maps the ONE shared
    host frame (rank 0 creates it, the name travels by broadcast_object_list) and stores only its own tile rows into it — what
    the blend kernels do over each rank's PCIe link; a fence; rank 0 reads the whole frame without any gather."""
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from houdini_gsplat_renderer_b200 import multigpu as M, scene as S
    from oracle import oracle as O
    cl = S.make_cloud(3000, 78, sh=False, scale_mult=2.5)
    fr = S.orbit_frame(200, 116, 10.0)
    cam = O.camera_from_view(fr.view)
    group = M.default_row_group(fr.height, world, bands_per_rank=2)
    name = [None]
    shm = None
    if rank == 0:
        shm = M.SharedHostFrames(fr.width, fr.height, 2)
        name[0] = shm.name
    dist.broadcast_object_list(name, src=0)
    if rank != 0:
        shm = M.SharedHostFrames(fr.width, fr.height, 2, name=name[0])
    ok = True
    for k in range(3):                                   # three frames: the two buffers alternate
        mine = O.pipeline(O.make_frame(fr, cam, cl.barycentre(), 0, row_rank=rank, row_world=world, row_group=group), cl)
        rows = M.owned_scanlines(fr.height, rank, world, group)
        dst = shm.frame(k)
        if rank == 0:
            dst[:] = -1.0                                # stale content of the buffer's previous use
        dist.barrier()
        dst[rows] = mine["rgba"][rows]                   # only the rows this rank owns
        dist.barrier()                                   # the frame fence
        if rank == 0:
            full = O.pipeline(O.make_frame(fr, cam, cl.barycentre(), 0), cl)
            ok = ok and bool(np.array_equal(dst, full["rgba"]))
        dist.barrier()
    if rank == 0:
        q.put(ok)
    shm.close()
    dist.barrier()
    dist.destroy_process_group()


def test_shared_host_frame_is_assembled_by_every_rank_writing_its_rows():
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
