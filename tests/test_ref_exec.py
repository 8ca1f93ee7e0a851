"""The parity pin: goldens produced by EXECUTING the unmodified reference sources
(/root/reference/{HashGrid,ParticleData,kernels/*,sesph,pcisph,iisph,dfsph}.py) under the serial
Taichi-semantics shim `oracle/tishim` (generator: tests/golden/make_ref_exec.py, run in the build
container; the .npz files are committed).  Scenes: 8^3 liquid block + boundary, 3 steps, every kernel
launch recorded.

CPU tests (this file, not gpu): the C oracle replays the launch stream FREE-RUNNING from the same initial
positions and must reproduce
  * HashGrid: gridCount, bucket contents, neighborCount and the ORDERED neighbour rows -- bit-exact;
  * every fp32 field every kernel wrote -- within FP_TOL (the oracle evaluates the same expressions in
    the same order; what is left is sqrt/divide grouping, a few ulp -- tolerances below);
  * the iteration counts of every convergence loop and the adapted time step.
GPU tests (-m gpu): the CUDA engine replays the same stream at the north star's 1e-4.

Documented differences to a real Taichi run (they are in the golden generator, not hidden here):
Q24/D-PCI (pcisph compute_nonpressure_force launched twice), out-of-bounds accesses Q7/Q12/Q15 read 0 /
are dropped (counted in meta["oob"]), serial ascending loop order where Taichi's atomics are unordered.
"""
import numpy as np
import pytest

from tests import refexec
from tests.refexec import Golden, OracleImpl, rel_err

# fp32 fields against the executed reference, scale-normalised max error max|a-b| / max|b|.
#   step 0: both sides start from bit-identical state -> what is compared is the arithmetic of every kernel: 1e-6
#           (measured: DFSPH bit-exact except the CG residual fields at 9e-7; SESPH rho / pressure bit-exact, d_vel 6e-8)
#   later : the oracle is FREE-RUNNING (nothing re-injected), so a last-ulp difference of step 0 is amplified by the stiff
#           terms (Tait pressure x7k, kappa / dt^2): 2e-5 over the recorded steps
FP_TOL_STEP0 = 2e-6
FP_TOL_FREE = 2e-5
# the PCG residual after an iteration is a cancelling difference (cg_r - alpha * A d is ~5 % of cg_r before the update), and alpha
# comes from two global sums: one ulp of the OLD residual shows up x20 relative to the NEW one
CANCELLING = {"cg_r": 10.0, "cg_s": 10.0, "cg_dir": 10.0}
ALL_GOLDENS = [(s, "") for s in refexec.SOLVERS] + [("sesph", "_kick"), ("pcisph", "_kick"), ("iisph", "_kick"), ("dfsph", "_kick"),
               ("dfsph", "_long")]        # _long: 6 free-running steps of a kicked 6^3 block, dv_iter 9..1, pr_iter 5..2, CFL-limited dt


# goldens the CPU oracle is held to but the CUDA replay is not run on (generated after the round's GPU budget was spent, so a GPU
# replay could not be validated): 6 / 5 free-running steps of kicked SESPH / PCISPH blocks
CPU_ONLY_GOLDENS = [("sesph", "_long"), ("pcisph", "_long")]


@pytest.mark.parametrize("solver,suffix", ALL_GOLDENS + CPU_ONLY_GOLDENS)
def test_oracle_reproduces_reference_executed_kernels(solver, suffix):
    g = Golden(solver, suffix)
    impl = OracleImpl(g)
    worst0, worst = {}, {}
    n_int = 0

    def check(idx, k, f, mine, gold):
        nonlocal n_int
        if f.startswith("hg_"):
            assert np.array_equal(np.asarray(mine), gold), "%s after %s (event %d, step %d): integer tables differ" % (f, k, idx, g.step_of(idx))
            n_int += 1
            return
        if f in refexec.GLOB:
            e = abs(mine - gold) / max(abs(gold), 1e-30) if gold != 0 else abs(mine)
        else:
            e = rel_err(mine, gold)
        w = worst0 if g.step_of(idx) == 0 else worst
        w[(k, f)] = max(w.get((k, f), 0.0), e / CANCELLING.get(f, 1.0))

    refexec.replay(g, impl, check)
    assert n_int >= 4
    bad = {k: v for k, v in worst0.items() if not v <= FP_TOL_STEP0}
    assert not bad, "step 0: oracle arithmetic departs from the executed reference: %s" % sorted(bad.items(), key=lambda kv: -kv[1])[:8]
    bad = {k: v for k, v in worst.items() if not v <= FP_TOL_FREE}
    assert not bad, "free-running oracle drifts from the executed reference: %s" % sorted(bad.items(), key=lambda kv: -kv[1])[:8]


@pytest.mark.parametrize("solver", refexec.SOLVERS)
def test_oracle_whole_steps_and_iteration_counts(solver):
    """oracle.step() (its own host loops, dfsph.py:84-164 etc.) K times == the reference's K frames"""
    g = Golden(solver)
    impl = OracleImpl(g)
    o = impl.o
    for s, info in enumerate(g.steps):
        o.step()
        for name in ("vs_iter", "dv_iter", "pr_iter"):
            if name in info:
                assert o.flag(name) == info[name], "step %d: %s %d != reference %d" % (s, name, o.flag(name), info[name])
        assert abs(o.get("deltaT") - info["deltaT"]) <= 1e-6 * info["deltaT"]
        assert rel_err(o.field("pos")[:g.nl], g.at_step_end(s, "pos")[:g.nl]) <= FP_TOL_FREE
        assert rel_err(o.field("vel"), g.at_step_end(s, "vel"), 1e-3) <= 5 * FP_TOL_FREE


@pytest.mark.parametrize("solver", refexec.SOLVERS)
def test_reference_executed_goldens_are_what_they_claim(solver):
    """provenance recorded in the fixture: commit, scene, the rewritten kernel list, the counted OOB events"""
    g = Golden(solver)
    m = g.meta
    assert m["reference_commit"] == "37f79c2" and m["solver"] == solver and len(g.steps) >= 3
    assert m["liquid_count"] == m["dim"] ** 2 * m.get("dimz", m["dim"]) and m["count"] == len(g.pos)
    assert "update_grid" in m["rewritten"]["HashGrid"] and "CubicGradW" in m["rewritten"]["kernels.CubicKernel"]
    oob = {(k, f, rw) for k, f, rw, n in m["oob"]}
    if solver == "dfsph":      # Q12: omega[j] / vel[j] with solid j (NL = 512 is a power of two: the max tree of Q15 stays in bounds;
        #                        the 8x8x7 "_kick" golden covers Q15)
        assert any(k == "compute_vorticity" and f in ("omega", "vel") and rw == "r" for k, f, rw in oob)
    if solver == "pcisph":     # Q7: rho_err[i] = 0 for i > 0
        assert any(f == "rho_err" and rw == "w" for _, f, rw in oob)


# ------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA engine against the executed reference
GPU_TOL = 1e-4          # the north star's figure: per-step fields within 1e-4 relative (scale-normalised)
# fields that are stiff images of other compared fields, with the scale the tolerance refers to
GPU_FLOORS = {
    ("sesph", "pressure"): 50000.0 * 7 * 0.1,       # Tait: 1e-4 x this = the pressure image of a 1e-5 relative density error
    ("pcisph", "pressure"): None,                   # filled per golden: delta / dt^2 x 1e-1 (see _gpu_floor)
}


def _gpu_floor(g, f):
    if (g.solver, f) == ("pcisph", "pressure"):
        dt = g.steps[0]["deltaT"]
        return float(g.meta["pci_coff"]) / (dt * dt) * 0.1
    if f in ("vel", "d_vel", "vel_guess", "cg_r", "cg_dir", "cg_Ad", "cg_s", "d_vel_pre", "vel_star", "omega", "d_omega", "normal", "dij_pj"):
        # velocity-like fields on a 1 cm/s scale; the PCG residual fields are velocity differences (cg_r = v - A v_guess) that shrink
        # towards 0 as the solve converges -- they are compared on the velocity scale, not on their own
        vmax = max([1e-2] + [float(np.abs(g.arr(i, "vel")).max()) for i, _, fs in g.events if "vel" in fs])
        return {"vel": 1e-2, "vel_guess": 1e-2, "vel_star": 1e-2, "omega": 1e-3, "d_omega": 1e-1,
                "cg_r": vmax, "cg_s": vmax, "cg_dir": vmax, "cg_Ad": vmax}.get(f, 0.0)
    return GPU_FLOORS.get((g.solver, f)) or 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("solver,suffix", ALL_GOLDENS)
def test_cuda_engine_matches_reference_executed_kernels(solver, suffix):
    """free-running replay of the reference's launch stream on the CUDA engine, every kernel's outputs at 1e-4;
    neighborCount (HashGrid.py:100) bit-exact; then whole fused steps must give the reference's iteration counts."""
    import torch
    assert torch.cuda.is_available()
    from tests.refexec import EngineImpl
    g = Golden(solver, suffix)
    kw = {"list_cap_liquid": 256, "list_cap_solid": 256}          # tiny scenes with dense 0.03-spaced walls: long solid lists
    impl = EngineImpl(g, **kw)
    worst = {}
    seen = set()
    cg0 = max([float(g.arr(i, "cg_delta_zero")[0]) for i, _, fs in g.events if "cg_delta_zero" in fs] + [0.0])

    def check(idx, k, f, mine, gold):
        if f == "hg_neighborCount":
            assert np.array_equal(np.asarray(mine), gold), "neighborCount differs after update_grid (event %d)" % idx
            seen.add(f)
            return
        if f in refexec.GLOB and gold == 0.0 and f != "deltaT":
            return      # Q18: a kernel that only ZEROES the accumulator inside its loop (iisph.py:322, dfsph.py:452); the engine clears it where it sums
        if f in refexec.GLOB:
            # the global sums are compared on the scale their loop test uses: avg_density_err / NL against 1e-3 (dfsph.py:160-163),
            # rho_err / NL against 1e-2 (pcisph.py:153-156), cg_delta against cg_delta_zero (dfsph.py:98)
            floor = {"avg_density_err": 1e-3 * g.nl, "rho_err": 1e-2 * g.nl, "cg_delta": cg0, "cg_delta_old": cg0}.get(f, 0.0)
            e = abs(mine - gold) / max(abs(gold), floor, 1e-30)
        else:
            e = rel_err(mine, gold, _gpu_floor(g, f))
        worst[(k, f)] = max(worst.get((k, f), 0.0), e)
        seen.add(f)

    refexec.replay(g, impl, check)
    assert impl.m.particle_data.hash_grid.status() == 0
    assert "hg_neighborCount" in seen and "pos" in seen
    bad = {k: v for k, v in worst.items() if not v <= GPU_TOL}
    assert not bad, "CUDA path departs from the executed reference: %s" % sorted(bad.items(), key=lambda kv: -kv[1])[:8]


@pytest.mark.gpu
@pytest.mark.parametrize("solver", refexec.SOLVERS)
def test_cuda_fused_steps_take_the_reference_iteration_counts(solver):
    import torch
    assert torch.cuda.is_available()
    from tests import util
    g = Golden(solver)
    m = util.make_engine(solver, g.pos, g.nl, list_cap_liquid=256, list_cap_solid=256)
    for s, info in enumerate(g.steps):
        m.step_fused(1)
        for name in ("vs_iter", "dv_iter", "pr_iter"):
            if name in info:
                assert getattr(m, name) == info[name], "step %d: %s %d != reference %d" % (s, name, getattr(m, name), info[name])
        dt = float(m.particle_data.deltaT.to_numpy()[0])
        assert abs(dt - info["deltaT"]) <= 1e-5 * info["deltaT"]          # CFL-limited: dt = 0.01 / sqrt(max |v + a dt|^2)
        assert rel_err(m.particle_data.pos.to_numpy()[:g.nl], g.at_step_end(s, "pos")[:g.nl]) <= GPU_TOL


# ------------------------------------------------------------------------------------------------------------------
# both branches of MarchingCubeGrid.export_surface, executed by the reference on a 4^3 scene after 2 steps
def _surface_setup():
    from oracle import oracle as orc
    g = Golden("dfsph", "_surface")
    impl = OracleImpl(g)
    end = g.steps[-1]["event_end"]
    seen = {}

    def check(idx, k, f, mine, gold):
        if not f.startswith("hg_") and f not in refexec.GLOB:
            seen[f] = max(seen.get(f, 0.0), rel_err(mine, gold))
    refexec.replay(g, impl, check, stop=end)
    assert seen["pos"] <= FP_TOL_FREE and seen["rho"] <= FP_TOL_FREE          # the two steps themselves (incl. the CFL-limited dt of step 2)
    o = impl.o
    # the surface kernels are compared as FUNCTIONS of the reference's state: same pos / rho in, same neighbour table
    o.field("pos")[...] = g.final("pos")
    o.field("rho")[...] = g.final("rho")
    tail = [e for e in g.events if e[0] >= end]
    mcm = g.meta["mc"]
    mc = orc.McOracle(g.pos, g.nl, 0.025, mcm["maxInGrid"], mcm["liqiudMass"], threads=1)
    assert list(mc.block) == mcm["block"] and np.allclose(mc.minb[0], mcm["min_boundary"], rtol=0, atol=0)
    return g, o, mc, tail, orc


def _next(it, name):
    for e in it:
        if e[1] == name:
            return e
    raise AssertionError("event %s missing" % name)


def test_oracle_surface_reconstruction_matches_reference_executed(golden_dir):
    """update_grid / cal_surface_point / marching_cube (MarchingCubeGrid.py:160-328): cell tables, colour field and the
    triangle soup BIT-EXACT; compute_color_map (ParticleData.py:188-218) bit-exact; cal_anistropic_kernel (:220-285) to 1e-6
    (the reference's ti.svd is LAPACK here, the oracle a Jacobi eigen-solver); the anisotropic colour field bit-exact from the
    reference's own G and to 1e-6 from the oracle's."""
    g, o, mc, tail, orc = _surface_setup()
    pos, rho = o.field("pos").copy(), o.field("rho").copy()
    it = iter(tail)
    mc.update_grid(pos)
    e = _next(it, "update_grid")
    gc, gg = g.arr(e[0], "mc_gridCount"), g.arr(e[0], "mc_grid")
    assert np.array_equal(mc.gridCount, gc)
    assert all(np.array_equal(mc.grid[c, :gc[c]], gg[c, :gc[c]]) for c in np.nonzero(gc)[0])
    sv = mc.cal_surface_point(rho).copy()
    e = _next(it, "cal_surface_point")
    assert np.array_equal(sv, g.arr(e[0], "mc_surface_value"))
    t = np.load("%s/mc_tables.npz" % golden_dir)
    n, tri = mc.marching_cube(t["edgetable"], t["tritable"])
    e = _next(it, "marching_cube")
    assert n == int(g.arr(e[0], "mc_vertex_count")[0]) and n > 1000
    assert np.array_equal(tri, g.arr(e[0], "mc_triangle"))
    color, grad = orc.compute_color_map(o)
    e = _next(it, "compute_color_map")
    assert np.array_equal(color, g.arr(e[0], "color")) and np.array_equal(grad, g.arr(e[0], "color_grad"))
    pa, G = orc.cal_anistropic_kernel(o, g.meta["mc"]["searchR"])
    e = _next(it, "cal_anistropic_kernel")
    gpa, gG = g.arr(e[0], "pos_avr"), g.arr(e[0], "G")
    assert rel_err(pa, gpa) <= 1e-6 and rel_err(G, gG) <= 1e-6
    mc.update_grid(pos)
    e = _next(it, "cal_surface_point_anistropic")
    gsv = g.arr(e[0], "mc_surface_value")
    assert np.array_equal(mc.cal_surface_point_anistropic(rho, gpa, gG), gsv)
    sva = mc.cal_surface_point_anistropic(rho, pa, G).copy()
    assert rel_err(sva, gsv) <= 1e-6
    n2, tri2 = mc.marching_cube(t["edgetable"], t["tritable"], surface_value=gsv)
    e = _next(it, "marching_cube")
    assert n2 == int(g.arr(e[0], "mc_vertex_count")[0]) and np.array_equal(tri2, g.arr(e[0], "mc_triangle"))
    # Q26 (found by executing the reference): cal_surface_point_anistropic reads G[j] / pos_avr[j] for SOLID j before its
    # `j < liquid_count` test (MarchingCubeGrid.py:229-238) -- out of bounds, value unused; counted in the fixture
    assert any(f in ("G", "pos_avr") and rw == "r" for _, f, rw, _n in g.meta["oob"])


# ------------------------------------------------------------------------------------------------------------------
# the canvas pass the scripts run every frame (Canvas.py:138-209 + draw_particle), executed by the reference
@pytest.mark.parametrize("solver", ["sesph", "dfsph"])
def test_oracle_canvas_matches_reference_executed(solver):
    from oracle import oracle as orc
    g = Golden(solver)
    sx = g.meta["img"]
    frames = [(i, k, f) for i, k, f in g.events if k == "draw_particle" and "canvas_img" in f]
    assert len(frames) >= 1          # a frame identical to the previous one is not stored again
    view = proj = None
    for idx, k, fields in g.events:
        if "canvas_view" in fields:
            view = g.arr(idx, "canvas_view")[0]
        if "canvas_proj" in fields:
            proj = g.arr(idx, "canvas_proj")[0]
        if k == "draw_particle" and "canvas_img" in fields:
            pos = g.at_step_end(g.step_of(idx), "pos")
            img, depth = orc.canvas_draw_particle(pos, g.nl, view, proj, sx, sx, 1 if solver == "dfsph" else 0)
            gi, gd = g.arr(idx, "canvas_img"), g.arr(idx, "canvas_depth")
            assert np.count_nonzero(gi) > 50
            assert np.array_equal(img, gi) and np.array_equal(depth, gd), "frame of step %d differs" % g.step_of(idx)


@pytest.mark.gpu
def test_cuda_surface_reconstruction_matches_reference_executed(golden_dir):
    """both branches of the surface reconstruction on the CUDA engine against what the reference itself computed
    (4^3 scene, state after 2 steps injected): active branch bit-exact, anisotropic branch within 1e-4."""
    import torch
    assert torch.cuda.is_available()
    from tests.refexec import EngineImpl
    g = Golden("dfsph", "_surface")
    impl = EngineImpl(g, list_cap_liquid=256, list_cap_solid=256)
    end = g.steps[-1]["event_end"]
    refexec.replay(g, impl, lambda *a: None, stop=end)
    m = impl.m
    pd = m.particle_data
    pd.pos.from_numpy(g.final("pos"))
    pd.rho.from_numpy(g.final("rho"))
    tail = [e for e in g.events if e[0] >= end]
    it = iter(tail)
    mc = pd.mc_grid
    assert [int(v) for v in mc.blocknp[0]] == g.meta["mc"]["block"]
    mc.update_grid()
    mc.cal_surface_point()
    e = _next(it, "cal_surface_point")
    assert np.array_equal(mc.surface_value.to_numpy(), g.arr(e[0], "mc_surface_value"))
    n = mc.marching_cube()
    e = _next(it, "marching_cube")
    assert n == int(g.arr(e[0], "mc_vertex_count")[0]) and np.array_equal(mc.mesh(), g.arr(e[0], "mc_triangle"))
    pd.compute_color_map()
    e = _next(it, "compute_color_map")
    assert rel_err(pd.color.to_numpy(), g.arr(e[0], "color")) <= GPU_TOL
    assert rel_err(pd.color_grad.to_numpy(), g.arr(e[0], "color_grad")) <= GPU_TOL
    pd.cal_anistropic_kernel()
    e = _next(it, "cal_anistropic_kernel")
    assert rel_err(pd.pos_avr.to_numpy(), g.arr(e[0], "pos_avr")) <= GPU_TOL
    assert rel_err(pd.G.to_numpy(), g.arr(e[0], "G")) <= GPU_TOL
    mc.update_grid()
    mc.cal_surface_point_anistropic()
    e = _next(it, "cal_surface_point_anistropic")
    gsv = g.arr(e[0], "mc_surface_value")
    assert rel_err(mc.surface_value.to_numpy(), gsv) <= GPU_TOL
    n2 = mc.marching_cube()
    e = _next(it, "marching_cube")
    n_ref = int(g.arr(e[0], "mc_vertex_count")[0])
    # topology: the same cubes are cut (a node within 1e-4 of the iso level may flip; none does on this scene)
    assert n2 == n_ref, (n2, n_ref)
    assert rel_err(mc.mesh(), g.arr(e[0], "mc_triangle")) <= 1e-3
    assert pd.hash_grid.status() == 0
