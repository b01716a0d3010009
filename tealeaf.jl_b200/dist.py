"""2-D domain decomposition: one process per GPU, one tile per process.

The reference has a single `Chunk` and no decomposition ("Parallelism strategies
in the reference: none"); its `haloupdate!` vocabulary ("exchange") comes from the MPI
upstream.  Here the global `xcells x ycells` interior is cut into a `px x py` grid of tiles
(y, the strided dimension, is split first so that most exchanged faces are contiguous rows).
Plumbing is `torch.distributed` (NCCL on GPUs, gloo in the CPU tests); the data path is
inside libtealeaf_b200: every tile maps the other tiles' memory (CUDA-IPC), the solver kernels
store their edge cells straight into the neighbours' halo cells over NVLink and sum the dot
products through peer-mapped mailboxes in their tails (option comm_fused=0 selects the older
halo-pull kernels + NCCL allreduces).
"""
from __future__ import annotations

import numpy as np

from .chunk import HostGeometry, paint_states
from .settings import Settings


class ThreadGroup:
    """`torch.distributed` look-alike for `world` tiles driven by `world` THREADS of one process
    (each thread calls `bind(rank)` first).  The tiles may live on different GPUs of the process
    or all on ONE GPU: their kernels run concurrently on per-tile streams and meet in the same
    peer-mapped mailboxes as on NVLink.  This is how the tiled code paths (halo pushes, mailbox
    sums, depth-k PPCG) are tested on a single-GPU box; it is a test harness, not a product mode
    (co-resident tiles share the GPU's bandwidth)."""
    same_process = True

    def __init__(self, world: int, grid=None):
        import threading
        self.world = world
        self.grid = grid
        self._barrier = threading.Barrier(world)
        self._slots = [None] * world
        self._local = threading.local()

    def bind(self, rank: int):
        self._local.rank = rank

    def get_rank(self):
        return self._local.rank

    def get_world_size(self):
        return self.world

    def barrier(self):
        self._barrier.wait()

    def all_gather_object(self, out, obj):
        self._slots[self.get_rank()] = obj
        self._barrier.wait()
        out[:] = list(self._slots)
        self._barrier.wait()

    def broadcast_object_list(self, lst, src=0):
        if self.get_rank() == src:
            self._slots[src] = list(lst)
        self._barrier.wait()
        lst[:] = list(self._slots[src])
        self._barrier.wait()

    def gather_object(self, obj, parts, dst=0):
        self._slots[self.get_rank()] = obj
        self._barrier.wait()
        if self.get_rank() == dst:
            parts[:] = list(self._slots)
        self._barrier.wait()

    def run(self, fn):
        """Run `fn(rank)` on one thread per tile; returns the list of results, re-raises the
        first exception (after breaking the barrier so the other threads do not wait forever)."""
        import threading
        results, errors = [None] * self.world, [None] * self.world

        def work(r):
            self.bind(r)
            try:
                results[r] = fn(r)
            except BaseException as exc:  # noqa: BLE001 - reported to the caller below
                errors[r] = exc
                self._barrier.abort()

        threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        real = [(r, e) for r, e in enumerate(errors) if e is not None and e.__class__.__name__ != "BrokenBarrierError"]
        if len(real) > 1:      # every tile's own story, not just the first one's
            raise RuntimeError("; ".join(f"[tile {r}] {e}" for r, e in real)) from real[0][1]
        if real or any(errors):
            raise (real[0][1] if real else [e for e in errors if e is not None][0])
        return results


def grid_for(world: int, dist=None):
    """px x py for `world` ranks: 1x1, 1x2, 2x2, 2x4, else the squarest split."""
    import os
    if dist is not None and getattr(dist, "grid", None):
        px, py = dist.grid
        if px * py != world:
            raise ValueError(f"grid {px}x{py} does not match {world} tiles")
        return px, py
    env = os.environ.get("TEALEAF_GRID")          # e.g. "1x4": override (tests, experiments)
    if env:
        px, py = (int(v) for v in env.lower().split("x"))
        if px * py != world:
            raise ValueError(f"TEALEAF_GRID={env} does not match {world} ranks")
        return px, py
    table = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
    if world in table:
        return table[world]
    px = int(np.floor(np.sqrt(world)))
    while world % px:
        px -= 1
    return px, world // px


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process (and hence its first-touch host allocations, pinned staging buffers
    included) to the CPUs of the NUMA node the GPU hangs off, so that host<->device copies of
    the 8 ranks of a box do not cross the socket interconnect.  Best effort: returns what it
    did, never raises."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            tok = visible.split(",")[device_index].strip()
            if tok.isdigit():
                idx = int(tok)
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"numa_node": node, "cpus": len(cpus)}
    except Exception as exc:  # no NVML / sysfs layout differs: leave the affinity alone
        info["error"] = str(exc)[:80]
    return info


def split(n: int, parts: int, idx: int):
    """(offset, size) of piece `idx` when n cells are cut into `parts` nearly equal pieces."""
    base, rem = divmod(n, parts)
    size = base + (1 if idx < rem else 0)
    off = idx * base + min(idx, rem)
    return off, size


def tile_of(rank: int, px: int, py: int, nx: int, ny: int):
    """(x0, y0, tile_nx, tile_ny) of `rank` = cx + cy*px."""
    cx, cy = rank % px, rank // px
    x0, tnx = split(nx, px, cx)
    y0, tny = split(ny, py, cy)
    return x0, y0, tnx, tny


def connect(chunk, dist):
    """Exchange CUDA-IPC blobs and the NCCL id, then wire the tile to its neighbours."""
    world = dist.get_world_size()
    blobs = [None] * world
    dist.all_gather_object(blobs, chunk.comm_export())
    if getattr(dist, "same_process", False):
        chunk.comm_connect(blobs, None)          # NCCL cannot put two ranks on one GPU; fused mode needs none
        return
    ident = [chunk.comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    chunk.comm_connect(blobs, ident[0])


def create_tile(settings: Settings, dist, device: int, backend=None, options=None):
    """`initialiseapp!` for this rank's tile.  Returns (chunk, geom, (px, py))."""
    from .app import upload_initial_state
    world, rank = dist.get_world_size(), dist.get_rank()
    px, py = grid_for(world, dist)
    x0, y0, tnx, tny = tile_of(rank, px, py, settings.xcells, settings.ycells)
    if backend is None:
        from .device import DeviceChunk
        backend = DeviceChunk
    chunk = backend(tnx, tny, settings.halodepth, settings.maxiters, device=device, rank=rank, px=px, py=py)
    connect(chunk, dist)
    for k, v in (options or {}).items():
        chunk.set_option(k, v)
    dist.barrier()       # every tile is wired and configured before the first rendezvous kernel is launched
    geom = HostGeometry(settings, tile=(x0, y0, tnx, tny))
    upload_initial_state(chunk, settings, geom)
    return chunk, geom, (px, py)


def gather_field(chunk, name: str, settings: Settings, dist, dst: int = 0):
    """Assemble the global (x, y) array of `name` (interior from every tile; halos from the
    tiles that own the physical boundary) on rank `dst`."""
    world, rank = dist.get_world_size(), dist.get_rank()
    px, py = grid_for(world, dist)
    local = chunk.get_field(name)
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst)
    if rank != dst:
        return None
    return assemble(parts, px, py, settings.xcells, settings.ycells, settings.halodepth)


def assemble(parts, px: int, py: int, nx: int, ny: int, hd: int) -> np.ndarray:
    out = np.zeros((nx + 2 * hd, ny + 2 * hd), order="F")
    for r, a in enumerate(parts):
        x0, y0, tnx, tny = tile_of(r, px, py, nx, ny)
        cx, cy = r % px, r // px
        # interior plus whichever halo sides are physical
        xl = 0 if cx == 0 else hd
        xr = tnx + 2 * hd if cx == px - 1 else tnx + hd
        yl = 0 if cy == 0 else hd
        yr = tny + 2 * hd if cy == py - 1 else tny + hd
        out[x0 + xl:x0 + xr, y0 + yl:y0 + yr] = a[xl:xr, yl:yr]
    return out


def paint_global_from_tiles(settings: Settings, px: int, py: int):
    """Host-only helper (tests): paint every tile and assemble; must equal the 1-tile painting."""
    dens, ener = [], []
    for r in range(px * py):
        g = HostGeometry(settings, tile=tile_of(r, px, py, settings.xcells, settings.ycells))
        d, e, _ = paint_states(settings, g)
        dens.append(d)
        ener.append(e)
    hd = settings.halodepth
    return (assemble(dens, px, py, settings.xcells, settings.ycells, hd),
            assemble(ener, px, py, settings.xcells, settings.ycells, hd))
