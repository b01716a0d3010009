"""Tiled (2-D domain decomposition) solves on ONE GPU: N tiles driven by N threads of this process
(`tealeaf_jl_b200.dist.ThreadGroup`), all on cuda:0.  Their kernels run concurrently on per-tile
streams and use exactly the multi-GPU data path -- edge cells stored into the neighbour tiles'
halo cells, dot products summed through the peer-mapped mailboxes in the kernel tails, depth-k
PPCG groups -- so the decomposition logic is covered by the single-GPU `-m gpu` run as well
(tests/test_multi_gpu.py repeats it across real GPUs when the box has several).

Every case is checked against the single-chunk CPU oracle with the north-star bar: identical
iteration counts (CG: +-1), per-step summaries within 1e-10, u / energy within 1e-9."""
import numpy as np
import pytest

import tealeaf_jl_b200 as tl
from tealeaf_jl_b200 import dist as tld
from conftest import classic_settings

pytestmark = pytest.mark.gpu


def run_tiled(grid, solver, nx, ny, steps=1, over=None, options=None, fields=("u", "energy")):
    """Solve the classic deck tiled `grid` = (px, py) on cuda:0.  Returns rank 0's
    (records, per-step summaries, {field: assembled global array})."""
    over = over or {}
    world = grid[0] * grid[1]
    tg = tld.ThreadGroup(world, grid)

    def fn(rank):
        s = classic_settings(nx, ny=ny, steps=steps, solver=solver, **over)
        chunk, geom, _ = tld.create_tile(s, tg, 0, options=options)
        try:
            summaries = []
            recs, final = tl.diffuse(chunk, s, geom,
                                     on_step=lambda rec: summaries.append(chunk.fieldsummary(geom.cell_volume)))
            out = {f: tld.gather_field(chunk, f, s, tg) for f in fields}
        finally:
            try:
                tg.barrier()      # nobody frees memory a neighbour's kernel may still address
            except Exception:
                pass
            chunk.close()
        return recs, summaries, out

    return tg.run(fn)[0]


def run_oracle(solver, nx, ny, steps=1, over=None):
    from oracle.oracle import OracleChunk
    s = classic_settings(nx, ny=ny, steps=steps, solver=solver, **(over or {}))
    oc, og = tl.initialiseapp(s, backend=OracleChunk)
    summaries = []
    recs, final = tl.diffuse(oc, s, og, on_step=lambda rec: summaries.append(oc.fieldsummary(og.cell_volume)))
    return recs, summaries, {"u": oc.get_field("u"), "energy": oc.get_field("energy")}


def check_against_oracle(got, ref, solver, hd=2):
    recs, sums, fields = got
    orecs, osums, ofields = ref
    its, oits = [r["iters"] for r in recs], [r["iters"] for r in orecs]
    slack = 1 if solver == "cg" else 0
    assert all(abs(a - b) <= slack for a, b in zip(its, oits)), (its, oits)
    serr = max(abs(a / b - 1) for sa, sb in zip(sums, osums) for a, b in zip(sa, sb))
    assert serr < 1e-10, serr
    for f in ("u", "energy"):
        a, b = fields[f][hd:-hd, hd:-hd], ofields[f][hd:-hd, hd:-hd]
        err = np.abs(a - b).max() / np.abs(b).max()
        assert err < 1e-9, (f, err)


CASES = [
    # grid, solver, nx, ny, settings overrides
    ((1, 2), "cg", 256, 192, {}),
    ((2, 1), "cg", 130, 77, {}),
    ((2, 2), "cg", 200, 150, {}),
    ((1, 2), "cheby", 192, 256, {}),
    ((2, 2), "cheby", 129, 67, {}),
    ((1, 2), "ppcg", 192, 160, {"ppcginnersteps": 6}),
    ((2, 2), "ppcg", 131, 150, {"ppcginnersteps": 5}),
    ((2, 2), "jacobi", 160, 130, {"maxiters": 120}),
    ((1, 4), "cg", 96, 256, {}),
    ((4, 1), "ppcg", 300, 64, {"ppcginnersteps": 4}),
]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("grid,solver,nx,ny,over", CASES, ids=[f"{g[0]}x{g[1]}-{s}-{nx}x{ny}" for g, s, nx, ny, _ in CASES])
def test_tiles_on_one_gpu_match_oracle(grid, solver, nx, ny, over):
    got = run_tiled(grid, solver, nx, ny, steps=1, over=over)
    check_against_oracle(got, run_oracle(solver, nx, ny, steps=1, over=over), solver)


@pytest.mark.timeout(600)
def test_two_timesteps_tiled():
    got = run_tiled((2, 2), "cg", 160, 160, steps=2)
    check_against_oracle(got, run_oracle("cg", 160, 160, steps=2), "cg")


# ---- matrix-powers PPCG: one tile exchange per k inner steps -------------------------------
DK_CASES = [
    # grid, nx, ny, inner steps, halo_depth, k
    ((1, 2), 192, 160, 6, 2, 2),
    ((2, 1), 150, 96, 5, 2, 2),
    ((2, 2), 131, 150, 5, 2, 2),
    ((2, 2), 150, 131, 10, 4, 4),
    ((2, 2), 129, 140, 7, 4, 3),
    ((1, 4), 96, 256, 8, 3, 3),
    ((4, 1), 300, 64, 4, 4, 4),
    ((3, 3), 200, 190, 9, 5, 5),
    ((2, 2), 128, 128, 3, 8, 8),      # k larger than the number of inner steps
]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("grid,nx,ny,inner,hd,k", DK_CASES,
                         ids=[f"{g[0]}x{g[1]}-{nx}x{ny}-inner{n}-hd{hd}-k{k}" for g, nx, ny, n, hd, k in DK_CASES])
def test_ppcg_depth_k_is_bit_identical_to_depth_1(grid, nx, ny, inner, hd, k):
    """Grouping the inner steps (halo_depth_k = k) must not change a single bit relative to the
    exchange-every-step schedule: the cells a tile computes in its halos are the owner's
    arithmetic on the owner's inputs.  Both are also held against the oracle."""
    over = {"ppcginnersteps": inner, "halodepth": hd}
    fields = ("u", "energy", "r", "sd", "p")
    base = run_tiled(grid, "ppcg", nx, ny, over={**over, "ppcghalodepth": 1}, fields=fields)
    deep = run_tiled(grid, "ppcg", nx, ny, over={**over, "ppcghalodepth": k}, fields=fields)
    assert [r["halo_depth_k"] for r in base[0]] == [1] and [r["halo_depth_k"] for r in deep[0]] == [min(k, inner)]
    assert [r["iters"] for r in base[0]] == [r["iters"] for r in deep[0]]
    assert [r["error"] for r in base[0]] == [r["error"] for r in deep[0]]
    assert base[0][0]["inner_total"] > 0
    for f in fields:
        a, b = base[2][f][hd:-hd, hd:-hd], deep[2][f][hd:-hd, hd:-hd]
        np.testing.assert_array_equal(a, b, err_msg=f)
    assert base[1] == deep[1]
    check_against_oracle(deep, run_oracle("ppcg", nx, ny, over=over), "ppcg", hd=hd)


@pytest.mark.timeout(600)
def test_ppcg_default_schedule_two_timesteps():
    """halo_depth_k = 0 (automatic): two inner steps per pass on the tiles (an exchange every 2 steps); with the pair
    kernels switched off, matrix-powers groups of halo_depth steps.  State carried across timesteps."""
    over = {"ppcginnersteps": 6, "halodepth": 3}
    ref = run_oracle("ppcg", 140, 120, steps=2, over=over)
    got = run_tiled((2, 2), "ppcg", 140, 120, steps=2, over=over)
    assert [r["halo_depth_k"] for r in got[0]] == [2, 2]
    check_against_oracle(got, ref, "ppcg", hd=3)
    got = run_tiled((2, 2), "ppcg", 140, 120, steps=2, over=over, options={"ppcg_pair": 0})
    assert [r["halo_depth_k"] for r in got[0]] == [3, 3]
    check_against_oracle(got, ref, "ppcg", hd=3)


# ---- temporal blocking on tiles: two Chebyshev iterations / PPCG inner steps per pass, depth-2 exchange ----------
PAIR_TILED_CASES = [((1, 2), 192, 256, 2), ((2, 1), 130, 77, 2), ((2, 2), 129, 67, 2), ((2, 2), 200, 150, 3), ((3, 3), 200, 190, 2),
                    ((1, 4), 96, 256, 2), ((4, 1), 301, 64, 2)]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("grid,nx,ny,hd", PAIR_TILED_CASES, ids=[f"{g[0]}x{g[1]}-{nx}x{ny}-hd{hd}" for g, nx, ny, hd in PAIR_TILED_CASES])
def test_tiled_cheby_pairs_are_bit_identical(grid, nx, ny, hd):
    """k_cheby_pair_ring<.., TILED>: every field bit-identical to one kernel (and one exchange) per iteration."""
    over = {"halodepth": hd}
    fields = ("u", "energy", "p", "w", "r")
    base = run_tiled(grid, "cheby", nx, ny, over=over, options={"pair_tiled": 0}, fields=fields)
    pair = run_tiled(grid, "cheby", nx, ny, over=over, fields=fields)      # default: pairs on tiles
    assert [(r["iters"], r["cg_iters"], r["cheby_iters"]) for r in base[0]] == [(r["iters"], r["cg_iters"], r["cheby_iters"]) for r in pair[0]]
    assert sum(r["kernel_launches"] for r in pair[0]) < sum(r["kernel_launches"] for r in base[0])
    for f in fields:
        np.testing.assert_array_equal(base[2][f][hd:-hd, hd:-hd], pair[2][f][hd:-hd, hd:-hd], err_msg=f)
    check_against_oracle(pair, run_oracle("cheby", nx, ny, over=over), "cheby", hd=hd)


PPCG_PAIR_TILED_CASES = [((1, 2), 192, 160, 2, 6), ((2, 1), 130, 77, 2, 5), ((2, 2), 129, 67, 2, 10), ((2, 2), 200, 150, 3, 7),
                         ((3, 3), 200, 190, 2, 4), ((1, 4), 96, 256, 2, 3), ((4, 1), 301, 64, 4, 2)]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("grid,nx,ny,hd,inner", PPCG_PAIR_TILED_CASES,
                         ids=[f"{g[0]}x{g[1]}-{nx}x{ny}-hd{hd}-inner{n}" for g, nx, ny, hd, n in PPCG_PAIR_TILED_CASES])
def test_tiled_ppcg_pairs_match(grid, nx, ny, hd, inner):
    """k_ppcg_pair_ring<.., TILED> (+ the trailing single step of an odd count): same iteration counts as one kernel and
    one exchange per inner step; fields agree to the rounding of the outer dot products (summed over different warp
    tasks), and both match the oracle."""
    over = {"halodepth": hd, "ppcginnersteps": inner}
    fields = ("u", "energy", "p", "sd", "r")
    base = run_tiled(grid, "ppcg", nx, ny, over={**over, "ppcghalodepth": 1}, fields=fields)
    pair = run_tiled(grid, "ppcg", nx, ny, over=over, fields=fields)
    assert [r["halo_depth_k"] for r in base[0]] == [1] and [r["halo_depth_k"] for r in pair[0]] == [2]
    key = lambda recs: [(r["iters"], r["cg_iters"], r["cheby_iters"], r["inner_total"]) for r in recs]
    assert key(base[0]) == key(pair[0])
    assert sum(r["kernel_launches"] for r in pair[0]) < sum(r["kernel_launches"] for r in base[0])
    scale = np.abs(base[2]["u"]).max()
    for f in fields:
        assert np.abs(base[2][f][hd:-hd, hd:-hd] - pair[2][f][hd:-hd, hd:-hd]).max() <= 1e-11 * scale, f
    check_against_oracle(pair, run_oracle("ppcg", nx, ny, over=over), "ppcg", hd=hd)


# ---- split tile exchange: post in the kernel tail, collect at the next kernel's entry (option xchg_deferred) -----
SPLIT_CASES = [((2, 2), "cg", 200, 150, {}), ((1, 4), "cg", 96, 256, {}), ((2, 2), "cheby", 129, 140, {}), ((3, 2), "cheby", 200, 131, {"halodepth": 3}),
               ((2, 2), "ppcg", 131, 150, {"ppcginnersteps": 5}), ((2, 2), "ppcg", 160, 120, {"ppcginnersteps": 6, "ppcghalodepth": 2}),
               ((2, 1), "ppcg", 130, 77, {"ppcginnersteps": 4, "ppcghalodepth": 1}), ((2, 2), "jacobi", 160, 130, {"maxiters": 120}),
               ((2, 2), "cheby", 150, 131, {"maxiters": 57})]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("grid,solver,nx,ny,over", SPLIT_CASES, ids=[f"{g[0]}x{g[1]}-{s}-{nx}x{ny}-{'-'.join(f'{k}{v}' for k, v in o.items())}" for g, s, nx, ny, o in SPLIT_CASES])
def test_split_exchange_is_bit_identical(grid, solver, nx, ny, over):
    """xchg_deferred = 1: the same packets, the same rank-order sums -- every field, iteration count and error value is
    bit-identical to the blocking exchange, two timesteps (state carried over), and matches the oracle."""
    fields = ("u", "energy", "p", "r", "sd")
    hd = over.get("halodepth", 2)
    base = run_tiled(grid, solver, nx, ny, steps=2, over=over, options={"xchg_deferred": 0}, fields=fields)
    split = run_tiled(grid, solver, nx, ny, steps=2, over=over, options={"xchg_deferred": 1}, fields=fields)
    key = lambda recs: [(r["iters"], r["cg_iters"], r["cheby_iters"], r["inner_total"], r["error"]) for r in recs]
    assert key(base[0]) == key(split[0])
    assert base[1] == split[1]
    for f in fields:
        np.testing.assert_array_equal(base[2][f][hd:-hd, hd:-hd], split[2][f][hd:-hd, hd:-hd], err_msg=f)
    check_against_oracle(split, run_oracle(solver, nx, ny, steps=2, over=over), solver, hd=hd)


# ---- lazy u update of CG kernel A (option cg_lazy_u, default on) on tiles ------------------------------------------
LAZY_TILED_CASES = [((2, 2), "cg", 200, 150, {}), ((1, 4), "cg", 96, 256, {"maxiters": 41}), ((3, 2), "cg", 131, 77, {"maxiters": 40}),
                    ((2, 2), "cheby", 129, 140, {}), ((2, 2), "ppcg", 131, 150, {"ppcginnersteps": 5})]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("grid,solver,nx,ny,over", LAZY_TILED_CASES, ids=[f"{g[0]}x{g[1]}-{s}-{nx}x{ny}" for g, s, nx, ny, o in LAZY_TILED_CASES])
def test_lazy_u_update_is_bit_identical_on_tiles(grid, solver, nx, ny, over):
    """u's tile-internal halos are filled after the loop (pull), so the lazy update needs nothing new on tiles: same bits."""
    fields = ("u", "energy", "p", "r", "w")
    base = run_tiled(grid, solver, nx, ny, steps=2, over=over, options={"cg_lazy_u": 0}, fields=fields)
    lazy = run_tiled(grid, solver, nx, ny, steps=2, over=over, options={"cg_lazy_u": 1}, fields=fields)
    key = lambda recs: [(r["iters"], r["cg_iters"], r["cheby_iters"], r["inner_total"], r["error"]) for r in recs]
    assert key(base[0]) == key(lazy[0])
    assert base[1] == lazy[1]
    for f in fields:
        np.testing.assert_array_equal(base[2][f], lazy[2][f], err_msg=f)
    check_against_oracle(lazy, run_oracle(solver, nx, ny, steps=2, over=over), solver)
