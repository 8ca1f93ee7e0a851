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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
