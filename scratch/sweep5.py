import sys, os, itertools
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import tealeaf_jl_b200 as tl
from conftest import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = classic_settings(N, steps=1, solver='cg', maxiters=400)
chunk, geom = tl.initialiseapp(s, backend=DeviceChunk)
rx, ry = s.dtinit / s.dx ** 2, s.dtinit / s.dy ** 2
def solve(tag):
    tl.haloupdate(chunk, s, 1, ["energy", "density"])
    chunk.copy_field("energy", "energy0")
    best = None
    for _ in range(3):
        info = chunk.cg_solve(s, rx, ry)
        m_ = info['solve_ms'] / info['iters']
        best = m_ if best is None else min(best, m_)
    print(f"N={N} {tag:60s} us/iter={best*1e3:.1f} Gcellit/s={N*N/best/1e6:.2f} err={info['error']:.6e}", flush=True)
def opts(**kw):
    for k, v in kw.items():
        chunk.set_option(k, v)
base = dict(hint_keep=0, hint_stream=0, b_reverse=0, l2_persist_mb=0)
opts(**base); solve('baseline')
for keep, stream, rev in itertools.product((0, 2), (0, 1), (0, 1)):
    if (keep, stream, rev) == (0, 0, 0): continue
    opts(hint_keep=keep, hint_stream=stream, b_reverse=rev); solve(f'keep={keep} stream={stream} rev={rev}')
opts(**base)
for mb in (32, 64, 96):
    for field in (6, 7):
        for hs in (1.0, 0.6):
            opts(l2_persist_field=field, l2_hit_scale=hs, l2_persist_mb=mb); solve(f'persist field={field} mb={mb} hit_scale={hs}')
opts(**base)
opts(hint_stream=1, b_reverse=1)
for mb in (32, 64):
    opts(l2_persist_field=6, l2_hit_scale=1.0, l2_persist_mb=mb); solve(f'stream=1 rev=1 + persist r mb={mb}')
opts(**base)
for pcr in (4, 8, 32):
    opts(hint_stream=1, b_reverse=1, pw_chunk_rows=pcr); solve(f'stream=1 rev=1 pw_chunk_rows={pcr}')
chunk.close()
