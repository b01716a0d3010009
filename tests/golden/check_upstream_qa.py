"""Runs the CPU oracle on the decks of tests/golden/upstream_qa.json and prints its temperature
next to upstream TeaLeaf's published QA value (the external pin of the oracle; see the JSON's
`provenance`).  Run from the repo root:

    python tests/golden/check_upstream_qa.py [--threads 8] [--min-cells 0] [--max-cells 2000] [--solver cg]

10..1000 cells take about a minute serially; 2000 x 2000 about two minutes on 8 threads.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import tealeaf_jl_b200 as tl  # noqa: E402
from conftest import classic_settings  # noqa: E402
from oracle.oracle import OracleChunk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--max-cells", type=int, default=1000)
    ap.add_argument("--min-cells", type=int, default=0)
    ap.add_argument("--solver", default="cg")
    a = ap.parse_args()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "upstream_qa.json")) as fh:
        cases = json.load(fh)["cases"]
    for c in cases:
        if c["x_cells"] > a.max_cells or c["x_cells"] < a.min_cells:
            continue
        s = classic_settings(c["x_cells"], ny=c["y_cells"], steps=c["end_step"], solver=a.solver)
        chunk, geom = tl.initialiseapp(s, backend=OracleChunk, threads=a.threads)
        t0 = time.time()
        recs, final = tl.diffuse(chunk, s, geom)
        print(f"{c['x_cells']:5d}^2  oracle {final['temp']!r:>22}  upstream {c['temp']!r:>22}  "
              f"rel {final['temp'] / c['temp'] - 1:+.2e}  iterations {sum(r['iters'] for r in recs)}  "
              f"{time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
