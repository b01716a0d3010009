# one Chebyshev and one PPCG solve at 4096^2 (for ncu captures of their kernels)
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
for solver in ('cheby', 'ppcg'):
    s = classic_settings(4096, steps=1, solver=solver, maxiters=1100)
    chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
    recs, final = tl.diffuse(chunk, s, geom)
    print(solver, recs[0], flush=True)
    chunk.close()
