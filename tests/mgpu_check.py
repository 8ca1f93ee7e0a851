"""Multi-GPU parity check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_check.py [nx ny nz steps [solver]]
Every rank runs the z-slab engine; rank 0 also runs the CPU oracle on the whole scene and compares
iteration counts and the gathered fields (same tolerance as the single-GPU parity tests)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcsph_b200 import dfsph, scenes  # noqa: E402


def main_other(solver, rank, world, nx, ny, nz, steps):
    """SESPH / IISPH / PCISPH on z-slab ranks, free running against the oracle: positions, densities (and pressure where it
    is not a stiff image of the density rounding), iteration counts, neighborCount exact."""
    import importlib
    mod = importlib.import_module("wcsph_b200." + solver)
    pts, nl = scenes.dam_break(nx, ny, nz, jitter=True, config_id=1)
    mod.init_scene(pts, nl, world_size=world, rank=rank)
    mod.reset_param()
    pd = mod.particle_data
    ok = True
    o = None
    if rank == 0:
        from oracle.oracle import Oracle
        o = Oracle(solver, pts, nl, threads=8)
    p_floor = {"sesph": 50000.0 * 7e-5, "iisph": 1.0}.get(solver)
    for s in range(steps):
        mod.step_fused(1)
        its = (getattr(mod, "vs_iter", 0), getattr(mod, "pr_iter", 0))
        pos, rho, prs = pd.pos.to_numpy(), pd.rho.to_numpy(), pd.pressure.to_numpy()
        nc = pd.hash_grid.neighborCount.to_numpy()
        flags = pd.hash_grid.status()
        if rank == 0:
            o.step()
            ito = (o.flag("vs_iter") if solver == "iisph" else 0, o.flag("pr_iter") if solver != "sesph" else 0)
            e_pos = np.abs(pos - o.field("pos")).max() / np.abs(o.field("pos")).max()
            e_rho = np.abs(rho - o.field("rho")).max() / np.abs(o.field("rho")).max()
            e_prs = 0.0 if p_floor is None else np.abs(prs - o.field("pressure")).max() / max(np.abs(o.field("pressure")).max(), p_floor)
            nc_ok = np.array_equal(nc, o.field("neighborCount"))
            good = its == ito and e_pos <= 1e-4 and e_rho <= 1e-4 and e_prs <= 1e-3 and nc_ok and flags == 0
            ok = ok and good
            print("%s step %d iters %s oracle %s err pos %.2e rho %.2e pressure %.2e neighborCount %s flags %d %s" % (
                solver, s, its, ito, e_pos, e_rho, e_prs, "exact" if nc_ok else "DIFFERS", flags, "ok" if good else "FAIL"), flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if t.item() == 1 else "FAIL")
    sys.exit(0 if t.item() == 1 else 1)


def kick_velocity(pts, nl, amp=2.0):
    """deterministic initial velocity of the moving-scene check: a +z drift (so particles cross every slab face within a few
    steps) with an x / y shear on top (so the viscosity and divergence solves have real work).  The drift is kept below the 5 cm
    gap to the +z wall over the checked steps: a particle that leaves the initial bounding box loses all its neighbours
    (HashGrid.py:81), and WHEN it crosses flips with the last bit of its position -- parity is not defined across that event."""
    p = np.asarray(pts[:nl], dtype=np.float64)
    # (the sin(6x), sin(8y), cos(7z) terms make the field compressive: the divergence solver has to iterate)
    v = np.stack([0.4 * np.sin(7.0 * p[:, 2]) + 0.3 * np.sin(6.0 * p[:, 0]), 0.3 * np.cos(5.0 * p[:, 0]) + 0.2 * np.sin(8.0 * p[:, 1]),
                  0.5 + 0.12 * np.sin(9.0 * p[:, 1]) + 0.1 * np.cos(7.0 * p[:, 2])], axis=1) * amp
    return v.astype(np.float32)


def slab_parity(world, rank, nx=12, ny=12, nz_per_rank=8, steps=12, amp=3.0, verbose=False):
    """DFSPH on `world` z-slab ranks, scene IN MOTION, free running against the CPU oracle (rank 0 runs it): iteration counts of
    all three loops equal per step, neighborCount exact, rho / pos within 1e-4, and particles really migrate across every
    interior face.  torch.distributed must be initialised (nccl).  Returns the summary dict on every rank."""
    import ctypes as C
    from wcsph_b200 import _lib
    nz = nz_per_rank * world + 16
    pts, nl = scenes.dam_break(nx, ny, nz, jitter=True, config_id=5)
    import importlib
    mod = importlib.reload(dfsph)
    mod.init_scene(pts, nl, world_size=world, rank=rank)
    mod.reset_param()
    pd = mod.particle_data
    v0 = kick_velocity(pts, nl, amp)
    pd.vel.from_numpy(v0)
    o = None
    if rank == 0:
        from oracle.oracle import Oracle
        o = Oracle("dfsph", pts, nl, threads=os.cpu_count() or 8)
        o.field("vel")[...] = v0
    worst = {"pos": 0.0, "rho": 0.0}
    q999 = {"pos": 0.0, "rho": 0.0}
    q99 = {"pos": 0.0, "rho": 0.0}
    outliers = 0
    iters_equal, nc_exact, flags_all = True, True, 0
    its_max = [0, 0, 0]
    for s in range(steps):
        mod.step_fused(1)
        it = (mod.vs_iter, mod.dv_iter, mod.pr_iter)
        pos, rho = pd.pos.to_numpy(), pd.rho.to_numpy()
        nc = pd.hash_grid.neighborCount.to_numpy()
        flags_all |= pd.hash_grid.status()
        # Field.to_numpy() already assembles the global field on slab ranks (rows of other ranks read 0, summed over the ranks)
        if rank == 0:
            o.step()
            ito = (o.flag("vs_iter"), o.flag("dv_iter"), o.flag("pr_iter"))
            iters_equal = iters_equal and it == ito
            its_max = [max(a, b) for a, b in zip(its_max, it)]
            op, orh = o.field("pos"), o.field("rho")
            ep = np.abs(pos[:nl] - op[:nl]).max(axis=1) / np.abs(op).max()
            er = np.abs(rho - orh) / np.abs(orh).max()
            worst["pos"] = max(worst["pos"], float(ep.max()))
            worst["rho"] = max(worst["rho"], float(er.max()))
            q999["pos"] = max(q999["pos"], float(np.quantile(ep, 0.999)))
            q999["rho"] = max(q999["rho"], float(np.quantile(er, 0.999)))
            q99["pos"] = max(q99["pos"], float(np.quantile(ep, 0.99)))
            q99["rho"] = max(q99["rho"], float(np.quantile(er, 0.99)))
            outliers = max(outliers, int(np.count_nonzero((ep > 1e-4) | (er > 1e-4))))
            nc_exact = nc_exact and np.array_equal(nc, o.field("neighborCount"))
            if verbose:
                print("slab step %d iters %s oracle %s err pos %.2e rho %.2e" % (s, it, ito, worst["pos"], worst["rho"]), flush=True)
    mc = (C.c_longlong * 5)()
    _lib.check(_lib.load().wcsph_migration_counts(pd._ctx, C.byref(mc)))
    mig = torch.tensor([mc[0], mc[1], mc[2], mc[3]], device="cuda", dtype=torch.int64)
    lst = [torch.zeros_like(mig) for _ in range(world)]
    if world > 1:
        dist.all_gather(lst, mig)
    else:
        lst = [mig]
    per_face_up = [int(lst[r][1].item()) for r in range(world - 1)]        # rank r -> r+1
    per_face_dn = [int(lst[r][0].item()) for r in range(1, world)]         # rank r -> r-1
    fl = torch.tensor([flags_all], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(fl, op=dist.ReduceOp.MAX)
    out = {"ranks": world, "scene": "dam_break(%d,%d,%d) jittered, kick %.1f m/s, %d steps" % (nx, ny, nz, amp, steps),
           "migrated_up_per_face": per_face_up, "migrated_down_per_face": per_face_dn,
           "migrated": int(sum(per_face_up) + sum(per_face_dn)),
           "max_rel_err": max(worst.values()), "err_pos": worst["pos"], "err_rho": worst["rho"],
           "p999_rel_err": max(q999.values()), "p99_rel_err": max(q99.values()), "particles_beyond_1e-4": outliers, "particles": int(nl),
           "iters_equal": bool(iters_equal), "iters_max_vs_dv_pr": its_max, "neighborCount_exact": bool(nc_exact),
           "status_flags": int(fl.item())}
    # The reference algorithm branches on `adv_rho[i] > 0` (dfsph.py:423) and `abs(sum) > eps` (:434): a particle whose Drho/Dt is
    # zero to rounding takes the warm-start correction in one implementation and not in the other.  From that step on the particle
    # carries a constant velocity difference (its position error grows linearly, 2e-5 per step) and the density of its whole
    # neighbourhood -- ~33 particles, 0.5 % of this small scene -- differs at the 1e-4 level; everything else stays at 1e-6
    # (dam_break(12,12,48): one such event at step 7, on ONE GPU and on 4 slab ranks alike, same 32 particles, same figures to four
    # digits).  Such events are counted, not averaged away: the check passes when 99 % of the particles stay within 1e-4 at every
    # step, at most 1 % (two flipped neighbourhoods) are beyond it, and nothing is beyond 1e-2.
    ok = (rank != 0) or (iters_equal and nc_exact and out["p99_rel_err"] <= 1e-4 and outliers <= max(2, nl // 100)
                         and out["max_rel_err"] <= 1e-2 and min(per_face_up + [1]) > 0 and out["status_flags"] == 0)
    t = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out["pass"] = bool(t.item() == 1)
    bl = [out]
    if world > 1:
        dist.broadcast_object_list(bl, src=0)
    return bl[0]


def main_checkpoint(rank, world):
    """SURVEY 8(f) N4 on slab ranks: save after 8 moving steps, run 5 more, restore (re-partitions by the restored positions), run
    the same 5 again -> same state as the uninterrupted run (the order inside a cell may differ after re-homing: 1e-5, not bit-exact)."""
    import tempfile
    from wcsph_b200 import checkpoint
    pts, nl = scenes.dam_break(12, 12, 8 * world + 16, jitter=True, config_id=5)
    dfsph.init_scene(pts, nl, world_size=world, rank=rank)
    dfsph.reset_param()
    pd = dfsph.particle_data
    pd.vel.from_numpy(kick_velocity(pts, nl, 2.0))
    dfsph.step_fused(8)
    path = [os.path.join(tempfile.gettempdir(), "wcsph_slab_ckpt_%d" % os.getpid()) if rank == 0 else None]
    dist.broadcast_object_list(path, src=0)
    checkpoint.save_state(dfsph, path[0])
    dist.barrier()
    dfsph.step_fused(5)
    ref_pos, ref_vel, ref_it = pd.pos.to_numpy(), pd.vel.to_numpy(), (dfsph.vs_iter, dfsph.dv_iter, dfsph.pr_iter)
    checkpoint.load_state(dfsph, path[0])
    dfsph.step_fused(5)
    pos, vel = pd.pos.to_numpy(), pd.vel.to_numpy()
    e_pos = float(np.abs(pos - ref_pos).max() / np.abs(ref_pos).max())
    e_vel = float(np.abs(vel - ref_vel).max() / max(np.abs(ref_vel).max(), 1e-2))
    ok = e_pos <= 1e-5 and e_vel <= 1e-4 and (dfsph.vs_iter, dfsph.dv_iter, dfsph.pr_iter) == ref_it and pd.hash_grid.status() == 0
    if rank == 0:
        print("checkpoint on %d slab ranks: err pos %.2e vel %.2e iters %s / %s" % (world, e_pos, e_vel, (dfsph.vs_iter, dfsph.dv_iter, dfsph.pr_iter), ref_it), flush=True)
        for ext in ("", ".npz"):
            if os.path.exists(path[0] + ext):
                os.unlink(path[0] + ext)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if t.item() == 1 else "FAIL")
    sys.exit(0 if t.item() == 1 else 1)


def main_surface(rank, world):
    """SURVEY 8(f) N2 on slab ranks: every rank evaluates the colour field of its own liquids, the shares are summed, the mesh of
    the total equals the restatement's (field to 1e-5: the summation order differs; same cubes cut)."""
    from oracle import oracle as orc
    pts, nl = scenes.dam_break(12, 12, 8 * world + 16, jitter=True, config_id=5)
    dfsph.init_scene(pts, nl, world_size=world, rank=rank)
    dfsph.reset_param()
    pd = dfsph.particle_data
    pd.vel.from_numpy(kick_velocity(pts, nl, 2.0))
    dfsph.step_fused(6)
    g = pd.mc_grid
    g.update_grid()
    g.cal_surface_point()
    nv = g.marching_cube()
    sv = g.surface_value.to_numpy()
    pos, rho = pd.pos.to_numpy(), pd.rho.to_numpy()
    ok = True
    if rank == 0:
        mo = orc.McOracle(pts, nl, 0.025, 4, pd.liqiudMass, threads=os.cpu_count() or 8)
        mo.update_grid(pos)
        ref = mo.cal_surface_point(rho).copy()
        t = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mc_tables.npz"))
        nref, _ = mo.marching_cube(t["edgetable"], t["tritable"])
        err = float(np.abs(sv - ref).max() / max(np.abs(ref).max(), 1e-30))
        ok = err <= 1e-5 and abs(nv - nref) <= max(6, nref // 500) and nv > 1000      # a node within 1e-6 of the iso value may flip a cube
        print("surface on %d slab ranks: field err %.2e, mesh vertices %d / %d" % (world, err, nv, nref), flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if t.item() == 1 else "FAIL")
    sys.exit(0 if t.item() == 1 else 1)


def main_moving(rank, world, args):
    nx, ny, nzr, steps = ([int(x) for x in args[:4]] + [12, 12, 8, 12][len(args[:4]):])
    res = slab_parity(world, rank, nx, ny, nzr, steps, verbose=(rank == 0))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB_PARITY", res)
        print("MGPU_CHECK", "PASS" if res["pass"] else "FAIL")
    sys.exit(0 if res["pass"] else 1)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if len(sys.argv) >= 2 and sys.argv[1] == "moving":
        return main_moving(rank, world, sys.argv[2:])
    if len(sys.argv) >= 2 and sys.argv[1] == "checkpoint":
        return main_checkpoint(rank, world)
    if len(sys.argv) >= 2 and sys.argv[1] == "surface":
        return main_surface(rank, world)
    a = [int(x) for x in sys.argv[1:5]] if len(sys.argv) >= 5 else [12, 12, 48, 12]
    nx, ny, nz, steps = a
    if len(sys.argv) >= 6 and sys.argv[5] in ("sesph", "iisph", "pcisph"):
        return main_other(sys.argv[5], rank, world, nx, ny, nz, steps)
    pts, nl = scenes.dam_break(nx, ny, nz, jitter=True, config_id=5)
    dfsph.init_scene(pts, nl, world_size=world, rank=rank)
    dfsph.reset_param()
    pd = dfsph.particle_data
    its = []
    ok = True
    o = None
    if rank == 0:
        from oracle.oracle import Oracle
        o = Oracle("dfsph", pts, nl, threads=8)
    for s in range(steps):
        dfsph.step_fused(1)
        its.append((dfsph.vs_iter, dfsph.dv_iter, dfsph.pr_iter))
        pos = pd.pos.to_numpy()
        rho = pd.rho.to_numpy()
        nc = pd.hash_grid.neighborCount.to_numpy()
        flags = pd.hash_grid.status()
        if rank == 0:
            o.step()
            ito = (o.flag("vs_iter"), o.flag("dv_iter"), o.flag("pr_iter"))
            e_pos = np.abs(pos - o.field("pos")).max() / np.abs(o.field("pos")).max()
            e_rho = np.abs(rho - o.field("rho")).max() / np.abs(o.field("rho")).max()
            nc_ok = np.array_equal(nc, o.field("neighborCount"))
            good = its[-1] == ito and e_pos <= 1e-4 and e_rho <= 1e-4 and nc_ok and flags == 0
            ok = ok and good
            print("step %d iters %s oracle %s err pos %.2e rho %.2e neighborCount %s flags %d %s" % (
                s, its[-1], ito, e_pos, e_rho, "exact" if nc_ok else "DIFFERS", flags, "ok" if good else "FAIL"), flush=True)
    # every rank splats its own slab, the per-pixel keys are min-reduced -> same picture as one GPU
    cv = dfsph.sph_canvas
    cv.static_cam(0.0, 1.0, 0.0)
    cv.clear_canvas()
    dfsph.draw_particle()
    img, depth = cv.img.to_numpy(), cv.depth.to_numpy()
    pos = pd.pos.to_numpy()
    if rank == 0:
        from oracle import oracle as _o
        oi, od = _o.canvas_draw_particle(pos, nl, cv.view[0], cv.proj[0], cv.sizex, cv.sizey, 1)
        good = np.array_equal(img, oi) and np.array_equal(depth, od)
        ok = ok and good
        print("canvas: %d lit pixels, %s" % (int(np.count_nonzero(img[:, :, 0])), "bit-exact" if good else "DIFFERS"), flush=True)
    import ctypes as C
    from wcsph_b200 import _lib
    n, gl, gh = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.load().wcsph_owned_count(pd._ctx, C.byref(n), C.byref(gl), C.byref(gh)))
    print("rank %d owns %d particles, ghosts %d / %d, slab z [%d, %d)" % (rank, n.value, gl.value, gh.value,
          pd.slab_plan["z_lo"][rank], pd.slab_plan["z_hi"][rank]), flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if t.item() == 1 else "FAIL")
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
