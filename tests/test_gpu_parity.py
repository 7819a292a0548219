"""GPU parity: libcspm_b200.so (through the C ABI) against the CPU oracle on
the same seeded inputs.  Bar: BIT-EXACT on every output — uint8 disparities,
f64 plane parameters, f64 aggregated costs (tolerance 0): the kernels execute
the reference's f64 operations in the reference's order without FMA
contraction, so there is nothing to tolerate."""
import numpy as np
import pytest

import orc
from crossscalepatchmatch_b200.synth import make_pair

pytestmark = pytest.mark.gpu

W, H, D = 96, 64, 24


@pytest.fixture(scope="module")
def cspm():
    import crossscalepatchmatch_b200 as c

    assert c.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return c


@pytest.fixture(scope="module")
def pair():
    l, r, gt = make_pair(W, H, D, 0xC5B20000)
    return l, r


def _queries(rng, w, h, d, n):
    x = rng.integers(0, w, n)
    y = rng.integers(0, h, n)
    # force border / corner pixels in
    x[:8] = [0, w - 1, 0, w - 1, 1, w - 2, w // 2, 0]
    y[:8] = [0, 0, h - 1, h - 1, 1, h - 2, 0, h // 2]
    v = rng.integers(0, 2, n)
    nm = rng.normal(size=(n, 3))
    nm /= np.linalg.norm(nm, axis=1)[:, None]
    nm[8:16] = [0, 0, 1]            # fronto-parallel
    nm[16:20, 2] = 1e-12            # |nz| < eps guard in update_param
    nm[20:24] = [0, 0, -1]          # negative nz
    z = rng.uniform(-2, d + 2, n)   # includes invalid disparities
    z[8:16] = np.arange(1, 9)       # integer disparities: fw == 1
    pt = np.stack([x, y, z], 1).astype(np.float64)
    return x, y, nm, pt, v


@pytest.mark.parametrize("S,lam", [(0, 0.0), (2, 0.3), (3, 0.3), (5, 0.3), (3, 0.0)])
def test_plane_cost_info_and_costs(cspm, orc_lib, pair, S, lam):
    l, r = pair
    cc = cspm.GetCCType("GRD")
    po = orc.PlaneCost(orc_lib, l, r, D, 35, S, lam)
    pg = cspm.PreSSPC(l, r, D, 35, cc) if S == 0 else cspm.PreCSPC(l, r, D, 35, S, cc, lam)
    wo, mo, do = po.info()
    wg, mg, dg = pg.info()
    assert np.array_equal(do, dg)
    assert np.array_equal(mo, mg), (mo, mg)
    assert np.array_equal(wo, wg), (wo, wg)
    for s in range(max(S, 1)):
        for v in (0, 1):
            assert np.array_equal(po.scale_image(s, v), pg.scale_image(s, v)), f"pyramid level {s} view {v}"
    x, y, nm, pt, v = _queries(np.random.default_rng(1), W, H, D, 4000)
    co = po.get_plane_cost(x, y, nm, pt, v)
    cg = pg.GetPlaneCostBatch(x, y, nm, pt, v)
    bad = np.flatnonzero(co != cg)
    assert bad.size == 0, f"{bad.size} costs differ, first {bad[:5]}: {co[bad[:5]]} vs {cg[bad[:5]]}"


def test_scalar_get_plane_cost(cspm, orc_lib, pair):
    l, r = pair
    pg = cspm.PreSSPC(l, r, D, 35, cspm.GetCCType("GRD"))
    po = orc.PlaneCost(orc_lib, l, r, D, 35, 0, 0.0)
    pl = cspm.Plane((0.0, 0.0, 1.0), (10.0, 12.0, 5.0))
    got = pg.GetPlaneCost(10, 12, pl, cspm.kLeft)
    want = po.get_plane_cost([10], [12], [[0, 0, 1.0]], [[10, 12, 5.0]], [0])[0]
    assert got == want


@pytest.mark.parametrize("right", [0, 1])
def test_build_cv(cspm, orc_lib, pair, right):
    l, r = pair
    lrgb = l[..., ::-1].astype(np.float64)
    rrgb = r[..., ::-1].astype(np.float64)
    cc = cspm.GrdCC()
    got = cc.buildRightCV(lrgb, rrgb, D + 1) if right else cc.buildCV(lrgb, rrgb, D + 1)
    want = orc_lib.build_cv("GRD", lrgb, rrgb, D + 1, right)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("right", [0, 1])
def test_build_cv_census(cspm, orc_lib, pair, right):
    l, r = pair
    rng = np.random.default_rng(3)
    # CV_64FC3 input with fractional values: convertTo(CV_8U) rounds half to even and saturates (cen_cc.cc:13)
    lrgb = l[..., ::-1].astype(np.float64) + rng.choice([0.0, 0.5, -0.5, 0.25], size=l.shape)
    rrgb = r[..., ::-1].astype(np.float64) + rng.choice([0.0, 0.5, 300.0, -7.0], size=r.shape, p=[0.9, 0.06, 0.02, 0.02])
    cc = cspm.CenCC()
    got = cc.buildRightCV(lrgb, rrgb, D + 1) if right else cc.buildCV(lrgb, rrgb, D + 1)
    want = orc_lib.build_cv("CEN", lrgb, rrgb, D + 1, right)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("S,lam", [(0, 0.0), (3, 0.3)])
def test_census_plane_cost_and_patchmatch(cspm, orc_lib, pair, S, lam):
    """cc_name=CEN (CenCC, cc/cen_cc.cc): the cost every input.txt line of the reference uses."""
    l, r = pair
    cc = cspm.GetCCType("CEN")
    po = orc.PlaneCost(orc_lib, l, r, D, 35, S, lam, cc_name="CEN")
    pg = cspm.PreSSPC(l, r, D, 35, cc) if S == 0 else cspm.PreCSPC(l, r, D, 35, S, cc, lam)
    for a, b in zip(po.info(), pg.info()):
        assert np.array_equal(a, b)
    assert po.info()[1].max() == 80.0
    x, y, nm, pt, v = _queries(np.random.default_rng(2), W, H, D, 3000)
    assert np.array_equal(po.get_plane_cost(x, y, nm, pt, v), pg.GetPlaneCostBatch(x, y, nm, pt, v))
    mo = orc.PatchMatch(orc_lib, l, r, D, 4, seed=21)
    mg = cspm.CSPatchMatch(l, r, D, 4, seed=21)
    mo.patchmatch(2, po, True)
    mg.PatchMatch(2, pg, True)
    _same_state(mo, mg, "census full run")
    for view in (0, 1):
        assert np.array_equal(mo.dis(view), mg.dis(view))


@pytest.mark.parametrize("S,lam", [(0, 0.0), (3, 0.3)])
def test_image_interpolating_plane_costs(cspm, orc_lib, pair, S, lam):
    """GrdPC (plane_cost/grd_pc.cc) and CSPC (plane_cost/cspc.cc): the alternative IPlaneCost plugins."""
    l, r = pair
    po = orc.PlaneCost(orc_lib, l, r, D, 35, S, lam, cc_name="IMG")
    pg = cspm.GrdPC(l, r, D, 35) if S == 0 else cspm.CSPC(l, r, D, 35, S, lam)
    for a, b in zip(po.info(), pg.info()):
        assert np.array_equal(a, b)
    x, y, nm, pt, v = _queries(np.random.default_rng(4), W, H, D, 3000)
    co, cg = po.get_plane_cost(x, y, nm, pt, v), pg.GetPlaneCostBatch(x, y, nm, pt, v)
    bad = np.flatnonzero(co != cg)
    assert bad.size == 0, f"{bad.size} costs differ, first {bad[:5]}: {co[bad[:5]]} vs {cg[bad[:5]]}"
    mo = orc.PatchMatch(orc_lib, l, r, D, 4, seed=33)
    mg = cspm.CSPatchMatch(l, r, D, 4, seed=33)
    mo.patchmatch(2, po, True)
    mg.PatchMatch(2, pg, True)
    _same_state(mo, mg, "GrdPC/CSPC full run")
    for view in (0, 1):
        assert np.array_equal(mo.dis(view), mg.dis(view))
    with pytest.raises(cspm.CspmError):
        cspm.GrdPC(l[:, :16].copy(), r[:, :16].copy(), 17, 35)  # max_disp > width: out-of-row reads in the reference


@pytest.mark.parametrize("wnd", [5, 11, 21, 41, 63])
def test_other_window_sizes(cspm, orc_lib, wnd):
    """wnd_size is a constructor argument of PreSSPC/PreCSPC (main.cc fixes 35): the generic-window
    code paths (runtime window size in every kernel, chunked pipeline in the spatial sweep)."""
    w, h, d = 70, 44, 14
    l, r, _ = make_pair(w, h, d, 500 + wnd)
    cc = cspm.GetCCType("GRD")
    po = orc.PlaneCost(orc_lib, l, r, d, wnd, 3, 0.3)
    pg = cspm.PreCSPC(l, r, d, wnd, 3, cc, 0.3)
    mo = orc.PatchMatch(orc_lib, l, r, d, 4, seed=5)
    mg = cspm.CSPatchMatch(l, r, d, 4, seed=5)
    mo.patchmatch(2, po, False)
    mg.PatchMatch(2, pg, False)
    _same_state(mo, mg, f"wnd={wnd}")


def _same_state(a, b, what):
    for view in (0, 1):
        for name, x, y in zip(("normal", "param", "min_cost"), a.planes(view), b.planes(view)):
            bad = np.argwhere(x != y)
            assert bad.size == 0, f"{what}: view {view} {name}: {len(bad)} mismatches, first at {bad[0]}: {x[tuple(bad[0])]} vs {y[tuple(bad[0])]}"


@pytest.mark.parametrize("S,lam", [(0, 0.0), (3, 0.3)])
def test_stages_bit_exact(cspm, orc_lib, pair, S, lam):
    """Drive the stages one by one; after each, GPU state must equal the oracle's
    bit for bit (and is re-synchronised from the oracle so a failure is local)."""
    l, r = pair
    cc = cspm.GetCCType("GRD")
    po = orc.PlaneCost(orc_lib, l, r, D, 35, S, lam)
    pg = cspm.PreSSPC(l, r, D, 35, cc) if S == 0 else cspm.PreCSPC(l, r, D, 35, S, cc, lam)
    mo = orc.PatchMatch(orc_lib, l, r, D, 4, seed=0x5EED)
    mg = cspm.CSPatchMatch(l, r, D, 4, seed=0x5EED)
    assert mg.n_ref() == orc_lib.pm_n_ref(mo.h_)
    mo.stage(orc.PatchMatch.INIT, 0, po)
    mg.stage(cspm.STAGE_INIT, 0, pg)
    _same_state(mo, mg, "init")
    for it in range(2):
        for so, sg, name in ((orc.PatchMatch.SPATIAL, cspm.STAGE_SPATIAL, "spatial"),
                             (orc.PatchMatch.VIEW, cspm.STAGE_VIEW, "view"),
                             (orc.PatchMatch.REFINE, cspm.STAGE_REFINE, "refine")):
            mo.stage(so, it, po)
            mg.stage(sg, it, pg)
            _same_state(mo, mg, f"iter {it} {name}")
    mo.stage(orc.PatchMatch.PLANE2DISP)
    mg.stage(cspm.STAGE_PLANE2DISP)
    for v in (0, 1):
        assert np.array_equal(mo.dis(v), mg.dis(v))
    mo.stage(orc.PatchMatch.POSTPROC, 0, po)
    mg.stage(cspm.STAGE_POSTPROC, 0, pg)
    for v in (0, 1):
        assert np.array_equal(mo.dis(v), mg.dis(v)), "post-processing"


@pytest.mark.parametrize("S,lam", [(0, 0.0), (3, 0.3)])
def test_spatial_shortcuts_are_exact_on_foreign_state(cspm, orc_lib, pair, S, lam):
    """The spatial sweep skips (a) a candidate that is bitwise the pixel's own plane and (b) a y-neighbour
    candidate equal to the x-neighbour's.  (a) is only sound while the stored costs are GetPlaneCost of the
    stored planes under the SAME plane cost; it must switch itself off for caller-supplied state
    (cspm_pm_set_planes) and when the plane cost changes between stages.  Large constant-plane regions make
    both shortcuts fire on (almost) every pixel."""
    l, r = pair
    cc = cspm.GetCCType("GRD")
    po = orc.PlaneCost(orc_lib, l, r, D, 35, S, lam)
    pg = cspm.PreSSPC(l, r, D, 35, cc) if S == 0 else cspm.PreCSPC(l, r, D, 35, S, cc, lam)
    mo = orc.PatchMatch(orc_lib, l, r, D, 4, seed=3)
    mg = cspm.CSPatchMatch(l, r, D, 4, seed=3)
    rng = np.random.default_rng(5)
    for v in (0, 1):
        n = np.zeros((H, W, 3)); n[..., 2] = 1.0            # fronto-parallel everywhere ...
        prm = np.zeros((H, W, 3)); prm[..., 2] = 5.0
        prm[:, W // 2:, 2] = 9.0                            # ... two constant regions
        cost = rng.uniform(0.0, 4000.0, (H, W))             # costs that do NOT belong to the planes
        mo.set_planes(v, n, prm, cost)
        mg.set_planes(v, n, prm, cost)
    for it in range(2):
        mo.stage(orc.PatchMatch.SPATIAL, it, po)
        mg.stage(cspm.STAGE_SPATIAL, it, pg)
        _same_state(mo, mg, f"foreign state, sweep {it}")
    # consistent state from one plane cost, then a sweep under another one
    po2 = orc.PlaneCost(orc_lib, l, r, D, 35, 2, 0.5)
    pg2 = cspm.PreCSPC(l, r, D, 35, 2, cc, 0.5)
    mo.stage(orc.PatchMatch.INIT, 0, po)
    mg.stage(cspm.STAGE_INIT, 0, pg)
    mo.stage(orc.PatchMatch.SPATIAL, 0, po)
    mg.stage(cspm.STAGE_SPATIAL, 0, pg)
    _same_state(mo, mg, "own state")
    mo.stage(orc.PatchMatch.SPATIAL, 1, po2)
    mg.stage(cspm.STAGE_SPATIAL, 1, pg2)
    _same_state(mo, mg, "plane cost swapped between stages")


@pytest.mark.parametrize("w,h,d,scale,S,lam,iters,pp", [
    (80, 60, 16, 4, 0, 0.0, 3, False),
    (80, 60, 16, 4, 3, 0.3, 2, True),
    (37, 41, 9, 8, 2, 1.0, 2, True),     # odd sizes, window larger than the image at coarse scales
    (130, 20, 40, 1, 0, 0.0, 1, False),  # wide, disparity range ~ image width: heavy wrap-around in view propagation
    (9, 7, 3, 2, 2, 0.3, 2, True),       # tiny images: every window is clipped on all sides
    (4, 3, 2, 2, 0, 0.0, 2, True),
    (1, 5, 1, 1, 0, 0.0, 2, True),       # single column
    (6, 1, 2, 1, 0, 0.0, 2, False),      # single row
    (2, 2, 1, 1, 0, 0.0, 1, False),
    (40, 33, 40, 1, 0, 0.0, 2, True),    # max_dis == width (largest legal range)
    (80, 60, 16, 4, 3, -0.2, 2, True),   # reg_lambda < 0 (main.cc:34 accepts it): a negative scale weight, sums not monotone
    (64, 48, 12, 4, 5, -0.3, 1, False),
])
def test_full_patchmatch_bit_exact(cspm, orc_lib, w, h, d, scale, S, lam, iters, pp):
    if min(w, h) >= 8 and d >= 4:
        l, r, _ = make_pair(w, h, d, 0xC5B20000 + w)
    else:  # degenerate sizes: plain random images
        rng = np.random.default_rng(w * 100 + h)
        l = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        r = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    cc = cspm.GetCCType("GRD")
    po = orc.PlaneCost(orc_lib, l, r, d, 35, S, lam)
    pg = cspm.PreSSPC(l, r, d, 35, cc) if S == 0 else cspm.PreCSPC(l, r, d, 35, S, cc, lam)
    mo = orc.PatchMatch(orc_lib, l, r, d, scale, seed=7)
    mg = cspm.CSPatchMatch(l, r, d, scale, seed=7)
    mo.patchmatch(iters, po, pp)
    mg.PatchMatch(iters, pg, pp)
    _same_state(mo, mg, "full run")
    for v in (0, 1):
        assert np.array_equal(mo.dis(v), mg.dis(v))
    st = mg.stats()
    evals_per_px = 1 + iters * (1 + mg.n_ref())
    assert st["plane_evals"] == 2 * w * h * evals_per_px + iters * 2 * ((w - 1) * h + w * (h - 1))
    assert st["kernel_launches"] > 0 and st["window_taps"] > 0


@pytest.mark.parametrize("env", [
    {"CSPM_SP_CTAS": "1"},                           # one CTA owns whole diagonals: up to 60 pixels, several passes of 32 tasks
    {"CSPM_SP_CTAS": "3"},                           # partition neighbours with cnt > 1 on both sides
    {"CSPM_WAVE_EXACT": "1"},                        # every task through the wavefront sweep's exact path
    {"CSPM_WAVE_EXACT": "1", "CSPM_SP_CTAS": "2"},
    {"CSPM_SPATIAL": "ring"},                        # round-1 sweep (serial add chains)
    {"CSPM_SPATIAL": "ring", "CSPM_SP_CTAS": "3"},   # ... with 8-20 pixels per CTA and multi-pass diagonals
    {"CSPM_SPATIAL": "diag"},                        # one launch per diagonal (fallback when a cooperative launch is refused)
], ids=lambda e: ",".join(f"{k[5:]}={v}" for k, v in e.items()))
def test_sweep_implementations_and_grid_caps(env):
    """The sweep implementations and grid sizes that the default run on a 148-SM GPU never takes at these image
    sizes: every CTA of the persistent grid owns at most one pixel of an 80x60 image, so the multi-pixel /
    multi-pass / neighbour-wait paths are forced here with CSPM_SP_CTAS, and the alternative sweeps with
    CSPM_SPATIAL / CSPM_WAVE_EXACT.  Each variant re-runs the oracle-checked stage and full-run tests."""
    import os
    import subprocess
    import sys

    if os.environ.get("CSPM_NESTED_TEST"):
        pytest.skip("nested run")
    here = os.path.abspath(__file__)
    sel = "test_stages_bit_exact or test_spatial_shortcuts or test_full_patchmatch_bit_exact or test_other_window_sizes or test_batch"
    out = subprocess.run([sys.executable, "-m", "pytest", here, "-x", "-q", "-m", "gpu", "-k", sel],
                         env={**os.environ, **env, "CSPM_NESTED_TEST": "1"}, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, f"{env}:\n{out.stdout[-3000:]}\n{out.stderr[-1500:]}"


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("S,lam", [(0, 0.0), (3, 0.3)])
def test_redblack_schedules_match_oracle(cspm, orc_lib, pair, schedule, S, lam):
    """CSPM_SCHEDULE_REDBLACK (4 neighbours) / _REDBLACK8 (+ the four at distance 5): checkerboard half-sweeps
    instead of the raster scan (include/cspm_c.h).  Not the reference's order — checked bit for bit against the
    oracle's restatement of the same schedule, stage by stage and as a full run; and it must differ from raster."""
    l, r = pair
    cc = cspm.GetCCType("GRD")
    po = orc.PlaneCost(orc_lib, l, r, D, 35, S, lam)
    pg = cspm.PreSSPC(l, r, D, 35, cc) if S == 0 else cspm.PreCSPC(l, r, D, 35, S, cc, lam)
    mo = orc.PatchMatch(orc_lib, l, r, D, 4, seed=11, schedule=schedule)
    mg = cspm.CSPatchMatch(l, r, D, 4, seed=11, schedule=schedule)
    mo.stage(orc.PatchMatch.INIT, 0, po)
    mg.stage(cspm.STAGE_INIT, 0, pg)
    for it in range(2):
        mo.stage(orc.PatchMatch.SPATIAL, it, po)
        mg.stage(cspm.STAGE_SPATIAL, it, pg)
        _same_state(mo, mg, f"red-black sweep {it}")
    mo = orc.PatchMatch(orc_lib, l, r, D, 4, seed=11, schedule=schedule)
    mg = cspm.CSPatchMatch(l, r, D, 4, seed=11, schedule=schedule)
    mo.patchmatch(2, po, True)
    mg.PatchMatch(2, pg, True)
    _same_state(mo, mg, "red-black full run")
    for v in (0, 1):
        assert np.array_equal(mo.dis(v), mg.dis(v))
    raster = cspm.CSPatchMatch(l, r, D, 4, seed=11)
    raster.PatchMatch(2, pg, True)
    assert not np.array_equal(raster.planes(0)[1], mg.planes(0)[1])


@pytest.mark.parametrize("S,lam,npairs", [(0, 0.0, 3), (3, 0.3, 2), (0, 0.0, 5)])
def test_batch_matches_individual_runs(cspm, orc_lib, S, lam, npairs):
    """cspm_pm_patchmatch_batch: several independent pairs whose spatial sweeps share one cooperative launch
    (lockstep diagonals, pair-major task list) must give exactly what pair-by-pair calls give — and the first
    pair is also checked against the oracle."""
    w, h, d = 80, 60, 16
    cc = cspm.GetCCType("GRD")
    pairs = [make_pair(w, h, d, 0xC5B20000 + 17 * k)[:2] for k in range(npairs)]
    mk_pc = lambda l, r: cspm.PreSSPC(l, r, d, 35, cc) if S == 0 else cspm.PreCSPC(l, r, d, 35, S, cc, lam)
    pcs = [mk_pc(l, r) for l, r in pairs]
    solo = [cspm.CSPatchMatch(l, r, d, 4, seed=100 + k) for k, (l, r) in enumerate(pairs)]
    for pm, pc in zip(solo, pcs):
        pm.PatchMatch(2, pc, True)
    batch = [cspm.CSPatchMatch(l, r, d, 4, seed=100 + k) for k, (l, r) in enumerate(pairs)]
    cspm.PatchMatchBatch(batch, pcs, 2, True)
    for k, (a, b) in enumerate(zip(solo, batch)):
        _same_state(a, b, f"pair {k} of the batch")
        for v in (0, 1):
            assert np.array_equal(a.dis(v), b.dis(v))
    l, r = pairs[0]
    po = orc.PlaneCost(orc_lib, l, r, d, 35, S, lam)
    mo = orc.PatchMatch(orc_lib, l, r, d, 4, seed=100)
    mo.patchmatch(2, po, True)
    _same_state(mo, batch[0], "batch pair 0 vs oracle")
    # mixed geometry is rejected
    other = cspm.CSPatchMatch(l[:50].copy(), r[:50].copy(), d, 4)
    with pytest.raises(cspm.CspmError):
        cspm.PatchMatchBatch([batch[0], other], [pcs[0], mk_pc(l[:50].copy(), r[:50].copy())], 1, False)


def test_error_behaviour(cspm, pair):
    l, r = pair
    assert cspm.GetCCType("BSM") is None            # main.cc:47-54
    with pytest.raises(cspm.CspmError):
        cspm.PreSSPC(l, r, D, 35, None)              # NULL cc_method
    with pytest.raises(cspm.CspmError):
        cspm.PreSSPC(l.astype(np.float64), r, D, 35, cspm.GetCCType("GRD"))  # CV_Assert(CV_8UC3)
    with pytest.raises(cspm.CspmError):
        cspm.PreCSPC(l, r, D, 35, 1, cspm.GetCCType("GRD"), 0.3)  # scale_num >= 2 (pre_cs_pc.cc:93)
    pg = cspm.PreSSPC(l, r, D, 35, cspm.GetCCType("GRD"))
    with pytest.raises(cspm.CspmError):
        pg.GetPlaneCostBatch([W], [0], [[0, 0, 1.0]], [[0, 0, 1.0]], [0])  # pixel outside the image
    with pytest.raises(cspm.CspmError):
        cspm.CSPatchMatch(l[:, :16].copy(), r[:, :16].copy(), 17, 1)  # max_dis > width: UB in the reference, rejected
    with pytest.raises(cspm.CspmError) as ei:
        cspm.CSPatchMatch(l, r, D, 4, schedule=7)     # unknown schedule
    assert ei.value.code == -1                       # CSPM_EINVAL travels through the NULL-returning creator
    small = cspm.CSPatchMatch(l[:32, :32].copy(), r[:32, :32].copy(), D, 4)
    with pytest.raises(cspm.CspmError):
        small.PatchMatch(1, pg)                      # geometry mismatch


def test_cpp_driver_matches_python_api(cspm, tmp_path):
    """apps/cspm_main (the reference's main() flow over include/cspm.hpp, C++ host
    code -> C ABI) must produce the same disparity maps as the Python mirror."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "apps", "cspm_main")
    assert os.path.exists(exe), "apps/cspm_main not built (run __graft_entry__.build())"
    w, h, d = 72, 50, 16
    l, r, _ = make_pair(w, h, d, 99)
    for name, img in (("l", l), ("r", r)):
        with open(tmp_path / f"{name}.ppm", "wb") as f:
            f.write(f"P6\n{w} {h}\n255\n".encode() + img[..., ::-1].tobytes())
    out = subprocess.run([exe, f"--l_img_file={tmp_path}/l.ppm", f"--r_img_file={tmp_path}/r.ppm",
                          f"--l_dis_file={tmp_path}/ld.pgm", f"--r_dis_file={tmp_path}/rd.pgm", "--max_dis=16",
                          "--dis_scale=4", "--cc_name=GRD", "--use_cs", "--reg_lambda=0.3", "--use_pp", "--seed=123"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "Total Time:" in out.stdout
    pg = cspm.PreCSPC(l, r, d, 35, 5, cspm.GetCCType("GRD"), 0.3)
    mg = cspm.CSPatchMatch(l, r, d, 4, seed=123)
    mg.PatchMatch(3, pg, True)
    for v, fn in ((0, "ld.pgm"), (1, "rd.pgm")):
        raw = open(tmp_path / fn, "rb").read()
        img = np.frombuffer(raw[-w * h:], np.uint8).reshape(h, w)
        assert np.array_equal(img, mg.dis(v))
    # error behaviour: unknown cost name -> failure exit, like the reference's NULL cc
    bad = subprocess.run([exe, f"--l_img_file={tmp_path}/l.ppm", f"--r_img_file={tmp_path}/r.ppm", "--max_dis=16",
                          "--dis_scale=4", "--cc_name=BSM"], capture_output=True, text=True)
    assert bad.returncode != 0
