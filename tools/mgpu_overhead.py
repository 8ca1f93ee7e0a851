"""step-time floor of the z-slab path: a scene so small that GPU work is negligible (launch, NCCL and host
sync latencies only).  torchrun --nproc-per-node R tools/mgpu_overhead.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcsph_b200 import dfsph, scenes, _lib
import ctypes as C
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ["NCCL_DEBUG"] = "NONE"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pts, nl = scenes.dam_break(12, 12, 16 * max(world, 2), jitter=True)
kw = dict(world_size=world, rank=rank) if world > 1 else {}
dfsph.init_scene(pts, nl, **kw)
dfsph.reset_param()
if world == 1 and len(sys.argv) > 1 and sys.argv[1] == "nograph":
    dfsph.set_graph(False)
dfsph.step_fused(5)
torch.cuda.synchronize()
K = 50
t0 = time.perf_counter()
dfsph.step_fused(K, fetch_iters=False)
t_enq = time.perf_counter() - t0
dfsph.particle_data.sync()
t_all = time.perf_counter() - t0
L = _lib.load(); ctx = dfsph.particle_data._ctx
_lib.check(L.wcsph_profile(ctx, 1))
dfsph.step_fused(K, fetch_iters=False)
buf = C.create_string_buffer(1 << 16)
_lib.check(L.wcsph_profile_report(ctx, buf, len(buf)))
if rank == 0:
    print("world %d NL %d: %.3f ms/step wall, host enqueue %.3f ms/step" % (world, nl, t_all / K * 1e3, t_enq / K * 1e3))
    rows = sorted(((float(t) / K, int(n) / K, nm) for nm, n, t in (l.split("\t") for l in buf.value.decode().splitlines())), reverse=True)
    print("  sum of event-timed regions %.3f ms/step" % sum(r[0] for r in rows))
    for t, n, nm in rows[:8]:
        print("   %-40s %5.1f/step %.4f ms/step" % (nm, n, t))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
