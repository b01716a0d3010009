# Stress of tl_create_multi with tiles sharing ONE GPU: create / initialise / short solve / destroy in a loop, after a
# warm-up of ordinary single-tile contexts (the flaky first rendezvous was only seen late in a long pytest process).
import os, sys, time
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.getcwd())
import tealeaf_jl_b200 as tl
from tealeaf_jl_b200.decks import classic_settings
from tealeaf_jl_b200.device import DeviceChunk
from tealeaf_jl_b200.lib import TeaLeafError
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for _ in range(12):      # the kind of traffic that precedes the multi tests in the suite
    s = classic_settings(200, ny=120, steps=1, solver="cg")
    c, g = tl.initialiseapp(s, backend=DeviceChunk); tl.diffuse(c, s, g); c.close()
fails = 0
for rep in range(reps):
    n = (4, 2, 6, 3)[rep % 4]
    s = classic_settings(200, ny=150, steps=1, solver=("cg", "cheby", "ppcg")[rep % 3], maxiters=150)
    t0 = time.time()
    try:
        c = DeviceChunk.multi(s.xcells, s.ycells, s.halodepth, s.maxiters, ngpus=n, devices=[0] * n)
    except TeaLeafError as e:
        print(f"[stress] rep {rep} n={n}: create failed: {e}", flush=True); fails += 1; continue
    ok_before = c.get_option("debug_comm_table_ok")
    try:
        g = tl.upload_initial_state(c, s)
        tl.diffuse(c, s, g)
        print(f"[stress] rep {rep} n={n} {s.solver}: ok ({time.time() - t0:.2f} s) table_ok={ok_before}", flush=True)
    except TeaLeafError as e:
        fails += 1
        try:
            ok_after = c.get_option("debug_comm_table_ok")
        except Exception as e2:
            ok_after = repr(e2)
        print(f"[stress] rep {rep} n={n} {s.solver}: FAILED table_ok before={ok_before} after={ok_after}: {str(e)[:700]}", flush=True)
    c.close()
print(f"[stress] {fails} failures in {reps} repetitions (TL_COMM_TABLE_LEGACY={os.environ.get('TL_COMM_TABLE_LEGACY')})")
