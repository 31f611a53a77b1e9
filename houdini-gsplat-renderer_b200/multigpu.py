"""Screen-tile partition across the GPUs of one box, one process per GPU.

Tiles are independent (a pixel depends only on the ordered list of splats covering it), so the path
shards by interleaved bands of tile rows: rank g owns tile rows ty with (ty // group) % G == g.  Every rank
holds all splat attributes, culls to its rows, selects / sorts / bins / blends only those.  The reference has
no multi-GPU path at all.

The one exchange step is the hand-off of finished tiles, and it is fused into the blend kernel:

* ``combine="p2p"``  – the display rank exports its frame with CUDA IPC, the other ranks map it and their blend
  kernels store every finished tile straight into it over NVLink peer memory (``gsb_target.final_rgba``).
* ``combine="host"`` – one frame in POSIX shared memory, page-locked and mapped by EVERY rank
  (``gsb_host_register``): each rank's blend kernel stores its finished tiles into it over the rank's own PCIe
  link, so an N-GPU frame reaches host memory over N links instead of one (no device gather, no staged D2H).
* ``combine="nccl"`` – the plain variant kept for comparison: every rank renders into its own frame, one NCCL
  reduction onto the display rank (exact, because every pixel is non-zero on exactly one rank: x + 0 == x).

A 4-byte NCCL all-reduce per frame is the completion fence in the first two modes.  Shared frames are double
buffered: frame i + 1 is stored into the other buffer, so the consumer of frame i (a D2H copy, the viewport) never
sees tiles of the next frame (the blend kernels of the other ranks start as soon as their own stream allows).
"""
from __future__ import annotations

import mmap
import os

import numpy as np

TILE = 16


def owned_tile_rows(height: int, rank: int, world: int, group: int = 1) -> list[int]:
    tiles_y = (height + TILE - 1) // TILE
    return [ty for ty in range(tiles_y) if (ty // max(group, 1)) % world == rank]


def owned_scanlines(height: int, rank: int, world: int, group: int = 1) -> np.ndarray:
    """bool[height]: scanlines (bottom-up, GL order) this rank renders."""
    y = np.arange(height)
    return (((y // TILE) // max(group, 1)) % world) == rank


def default_row_group(height: int, world: int, bands_per_rank: int = 8) -> int:
    """Band height in tile rows: wide enough that few splats straddle two ranks (each straddler is projected,
    sorted and binned twice), narrow enough that every rank still gets >= bands_per_rank bands for load balance."""
    tiles_y = (height + TILE - 1) // TILE
    return max(1, tiles_y // (max(world, 1) * bands_per_rank))


def combine_on_root(fb, rank: int, world: int, root: int = 0, group=None):
    """fb: [H,W,4] float32 torch tensor holding this rank's rows (zeros elsewhere).  After the call the
    root holds the whole frame.  NCCL on GPUs (NVLink 5 / NVSwitch), gloo in the CPU tests."""
    if world == 1:
        return fb
    import torch.distributed as dist
    dist.reduce(fb, dst=root, op=dist.ReduceOp.SUM, group=group)
    return fb


class SharedHostFrames:
    """``count`` RGBA32F frames in one POSIX shared-memory file (/dev/shm), mapped by every rank.  The creator (rank 0)
    passes ``name=None`` and publishes ``.name``; the others attach by name.  Page aligned, so it can be page-locked."""

    def __init__(self, width: int, height: int, count: int = 2, name: str | None = None):
        self.width, self.height, self.count = int(width), int(height), int(count)
        page = mmap.PAGESIZE
        self.frame_bytes = self.width * self.height * 16
        self.stride = (self.frame_bytes + page - 1) // page * page          # every frame starts on a page
        self.nbytes = self.stride * self.count
        self.owner = name is None
        self.name = name or f"gsb_frames_{os.getpid()}_{id(self) & 0xffff:x}"
        path = os.path.join("/dev/shm", self.name)
        flags = os.O_RDWR | (os.O_CREAT | os.O_EXCL if self.owner else 0)
        fd = os.open(path, flags, 0o600)
        try:
            if self.owner:
                os.ftruncate(fd, self.nbytes)
            self._mm = mmap.mmap(fd, self.nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        self.buffer = np.frombuffer(self._mm, dtype=np.uint8)

    def frame(self, k: int) -> np.ndarray:
        off = self.stride * (k % self.count)
        return self.buffer[off:off + self.frame_bytes].view(np.float32).reshape(self.height, self.width, 4)

    def close(self):
        self.buffer = None
        try:
            self._mm.close()
        except (BufferError, ValueError):
            pass                                 # a view is still alive somewhere: the mapping dies with the process
        if self.owner:
            try:
                os.unlink(os.path.join("/dev/shm", self.name))
            except FileNotFoundError:
                pass


class near_gpu:
    """Context manager: run the enclosed host-side setup on the CPUs NVML reports as local to CUDA device ``device`` (its
    NUMA node), then restore the previous affinity.  Page-locked host memory is placed where the thread that registers /
    first touches it runs; a frame that the blend kernels store into over PCIe should live on the GPU's own node.  A no-op
    when NVML, the device UUID or sched_setaffinity are unavailable (GSB_NUMA_AFFINITY=0 turns it off)."""

    def __init__(self, device: int):
        self.device, self.saved, self.cpus = device, None, None

    def __enter__(self):
        if os.environ.get("GSB_NUMA_AFFINITY", "1") == "0" or not hasattr(os, "sched_setaffinity"):
            return self
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.device).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus and cpus != allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus)
            self.cpus = sorted(cpus)
        except Exception:                                     # noqa: BLE001 - placement is an optimisation, never an error
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:                                 # noqa: BLE001
                pass
        return False


class RowPartitionedRenderer:
    """Per-rank driver of a row-partitioned frame: owns the shared frames, the per-frame fence and the hand-off.

    ``renderer`` is this rank's ``GSplatRenderer`` (already holding the replicated splats); ``dist`` is an initialised
    ``torch.distributed`` (NCCL).  ``render(frame)`` leaves the finished frame in the display rank's device frame
    (``device_frame()``); ``render(frame, to_host=True)`` leaves it in host memory on every rank (``host_frame()``), and
    returns once it is complete."""

    def __init__(self, renderer, width: int, height: int, rank: int, world: int, combine: str = "p2p",
                 row_group: int | None = None, stream=None):
        import torch
        import torch.distributed as dist
        assert combine in ("p2p", "nccl", "host")
        self.r, self.W, self.H, self.rank, self.world, self.combine = renderer, width, height, rank, world, combine
        self.row_group = default_row_group(height, world) if row_group is None else row_group
        self.torch, self.dist = torch, dist
        self.stream = stream or torch.cuda.current_stream()
        self.fence = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.k = 0                                            # frame counter: shared frames are double buffered
        self.peer = [None, None]                              # device pointers of the display rank's two frames
        self.local = None                                     # nccl mode: this rank's own frame
        self.last_host = None
        frame_bytes = width * height * 16
        if world > 1 and combine == "p2p":
            # two frames in ONE exported allocation on the display rank; the others map it (NVLink peer memory)
            hbuf = torch.zeros(64, dtype=torch.uint8, device="cuda")
            if rank == 0:
                handle, base = renderer.ipc_export_frame(width, height * 2)
                hbuf.copy_(torch.tensor(list(handle), dtype=torch.uint8))
            dist.broadcast(hbuf, 0)
            if rank != 0:
                base = renderer.ipc_open(bytes(hbuf.cpu().numpy().tobytes()))
            self._peer_base = base
            self.peer = [base, base + frame_bytes]
        if combine == "nccl":
            self.local = torch.zeros((height, width, 4), dtype=torch.float32, device="cuda")
        # host frames: one shared-memory file mapped by every rank, page-locked + device-mapped by each of them
        name = [None]
        self.shm = None
        with near_gpu(torch.cuda.current_device()) as ng:     # the frame's pages go to the GPU's own NUMA node
            self.host_cpus = ng.cpus
            if rank == 0:
                self.shm = SharedHostFrames(width, height, 2)
                name[0] = self.shm.name
            if world > 1:
                dist.broadcast_object_list(name, src=0)
            if rank != 0:
                self.shm = SharedHostFrames(width, height, 2, name=name[0])
            renderer.host_register(self.shm.buffer)

    def device_frame(self) -> int | None:
        """Display rank: device pointer of the frame the last render() completed (p2p / nccl modes)."""
        if self.world == 1:
            return self.r.device_framebuffer()
        if self.combine == "nccl":
            return self.local.data_ptr()
        return self.peer[(self.k - 1) & 1]

    def host_frame(self) -> np.ndarray | None:
        return self.last_host

    def render(self, frame, to_host: bool = False):
        r, world, rank, g = self.r, self.world, self.rank, self.row_group
        b = self.k & 1
        self.k += 1
        self.last_host = None
        if to_host:
            # every rank's blend stores its finished tiles into the shared pinned host frame over its own PCIe link
            host = self.shm.frame(b)
            r.render(frame, host_rgba=host, row_rank=rank, row_world=world, row_group=g)   # returns when this rank's rows are there
            if world > 1:
                self.dist.all_reduce(self.fence)              # frame fence: every rank's rows are in host memory
                self.stream.synchronize()
            self.last_host = host
        elif self.combine == "nccl" and world > 1:
            r.render(frame, device_rgba=self.local.data_ptr(), row_rank=rank, row_world=world, row_group=g)
            self.dist.reduce(self.local, dst=0, op=self.dist.ReduceOp.SUM)   # rows of the other ranks are zeros: exact
        elif world > 1:
            r.render(frame, final_rgba=self.peer[b], row_rank=rank, row_world=world, row_group=g)
            self.dist.all_reduce(self.fence)                  # frame fence (stream ordered after this rank's blend)
        else:
            r.render(frame)

    def close(self):
        r = self.r
        try:
            r.synchronize()
            if self.shm is not None:
                r.host_unregister(self.shm.buffer)
                self.shm.close()
            if self.rank != 0 and self.peer[0]:
                r.ipc_close(self._peer_base)
        finally:
            self.shm = None
