"""gloo world_size-2 worker (CPU): the host side of the tiled path -- decomposition, blob
all-gather / id broadcast as in dist.connect, gather_field assembly -- with a NumPy stand-in
for the device chunk (no compute: there is no CPU fallback of the product path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tealeaf_jl_b200 import dist as tld  # noqa: E402
from tealeaf_jl_b200.chunk import HostGeometry, paint_states  # noqa: E402
from conftest import classic_settings  # noqa: E402


class FakeChunk:
    def __init__(self, rank, fields):
        self.rank, self.fields, self.connected = rank, fields, None

    def comm_export(self):
        return bytes([self.rank]) * 16

    def comm_unique_id(self):
        return b"id-from-rank-%d" % self.rank

    def comm_connect(self, blobs, ident):
        self.connected = (blobs, ident)

    def get_field(self, name):
        return self.fields[name]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2
    s = classic_settings(40, ny=33)
    px, py = tld.grid_for(world)
    tile = tld.tile_of(rank, px, py, s.xcells, s.ycells)
    d, e, _ = paint_states(s, HostGeometry(s, tile=tile))
    chunk = FakeChunk(rank, {"density": d, "energy0": e})
    tld.connect(chunk, dist)
    blobs, ident = chunk.connected
    assert blobs == [bytes([0]) * 16, bytes([1]) * 16] and ident == b"id-from-rank-0"
    g = tld.gather_field(chunk, "density", s, dist)
    if rank == 0:
        dg, _, _ = paint_states(s, HostGeometry(s))
        np.testing.assert_array_equal(g, dg)
        print("dist_cpu_worker OK", flush=True)
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
