"""CPU coverage of the N > 1 host logic: the z-slab plan (pure numpy) and the rank-assembly rule of
Field.to_numpy (every rank fills only its rows, an all-reduce SUM assembles) under a world_size-2
gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wcsph_b200 import partition, scenes


def _plan(world):
    pts, nl = scenes.dam_break(10, 10, 40)
    f32 = pts.astype(np.float32)
    return pts, nl, partition.z_slabs(pts[:nl], f32.min(0), f32.max(0), 0.05, world)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_z_slabs_cover_grid_and_balance(world):
    pts, nl, plan = _plan(world)
    assert plan["z_lo"][0] == 0 and plan["z_hi"][-1] == plan["bz"]
    assert all(a == b for a, b in zip(plan["z_hi"][:-1], plan["z_lo"][1:]))          # contiguous, disjoint
    assert all(hi - lo >= 4 for lo, hi in zip(plan["z_lo"], plan["z_hi"]))           # >= 2 ghost layers per side
    assert sum(plan["counts"]) == nl
    assert max(plan["counts"]) - min(plan["counts"]) <= 2 * 10 * 10                   # within two z layers
    assert plan["cap_own"] >= max(plan["counts"]) and plan["cap_ghost"] >= 2 * 100


def test_z_slabs_rejects_thin_grid():
    pts, nl = scenes.dam_break(10, 10, 4)
    f32 = pts.astype(np.float32)
    with pytest.raises(ValueError):
        partition.z_slabs(pts[:nl], f32.min(0), f32.max(0), 0.05, 8)


def test_cell_z_matches_reference_arithmetic():
    # HashGrid.py:68: trunc(f32(pos - min) * f32(1/gridR)); lattice nodes sit exactly on cell faces (Q19)
    pts, nl = scenes.dam_break(4, 4, 12)
    f32 = pts.astype(np.float32)
    cz = partition.cell_z(pts[:nl], f32.min(0), 0.05)
    ref = ((f32[:nl, 2] - f32[:, 2].min()) * np.float32(1.0 / 0.05)).astype(np.int32)
    assert np.array_equal(cz, ref)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts, nl, plan = _plan(world)
    f32 = pts.astype(np.float32)
    cz = np.clip(partition.cell_z(pts[:nl], f32.min(0), 0.05), 0, plan["bz"] - 1)
    mine = (cz >= plan["z_lo"][rank]) & (cz < plan["z_hi"][rank])
    field = np.zeros((nl, 3), dtype=np.float32)
    field[mine] = f32[:nl][mine] * 2.0                     # what wcsph_field_get leaves on this rank
    t = torch.from_numpy(field)
    dist.all_reduce(t)                                     # Field.to_numpy(gather=True)
    cnt = torch.tensor([int(mine.sum())])
    dist.all_reduce(cnt)
    q.put((rank, bool(np.array_equal(t.numpy(), f32[:nl] * 2.0)), int(cnt.item()) == nl))
    dist.destroy_process_group()


def test_rank_assembly_under_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok1 and ok2 for _, ok1, ok2 in res), res


# ---- peer-mailbox negotiation (ParticleData._open_mailboxes): the decision is collective -------------------------------------------
class _FakeLib:
    """stands in for libwcsph_b200's mailbox entry points: rank `bad_rank` cannot map its peers"""

    def __init__(self, rank, bad_rank):
        self.rank, self.bad_rank, self.options, self.opened = rank, bad_rank, [], None

    def wcsph_comm_mailbox_handle(self, ctx, h):
        for k in range(64):
            h[k] = (self.rank * 7 + k) % 256
        return 0

    def wcsph_comm_mailbox_open(self, ctx, blob):
        self.opened = bytes(blob)
        return -2 if self.rank == self.bad_rank else 0

    def wcsph_set_option(self, ctx, name, value):
        self.options.append((name, value))
        return 0


def _mailbox_worker(rank, world, port, q, bad_rank, backend_name):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wcsph_b200.ParticleData import ParticleData
    pd = ParticleData.__new__(ParticleData)
    pd.world_size, pd.rank, pd.verbose = world, rank, False
    lib = _FakeLib(rank, bad_rank)

    class _Dist:                                 # gloo underneath, reporting the backend name under test
        def __getattr__(self, n):
            return getattr(dist, n)

        @staticmethod
        def get_backend():
            return backend_name
    # with the 'nccl' name the tensors would be moved to CUDA: keep them on the CPU for this host-logic test
    import torch as _t
    orig_cuda = _t.Tensor.cuda
    _t.Tensor.cuda = lambda self, *a, **k: self
    try:
        use = pd._open_mailboxes(lib, None, _Dist())
    finally:
        _t.Tensor.cuda = orig_cuda
    q.put((rank, use, lib.opened, lib.options))
    dist.destroy_process_group()


def _run_mailbox(bad_rank, backend_name, port_off):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() + port_off) % 500
    procs = [ctx.Process(target=_mailbox_worker, args=(r, 2, port, q, bad_rank, backend_name)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    return res


def test_mailbox_negotiation_all_ranks_open():
    res = _run_mailbox(bad_rank=-1, backend_name="nccl", port_off=11)
    expect = bytes((r * 7 + k) % 256 for r in range(2) for k in range(64))          # handles in rank order
    assert all(use for _, use, _, _ in res) and all(op == expect for _, _, op, _ in res) and all(not o for _, _, _, o in res)


def test_mailbox_negotiation_falls_back_collectively():
    """one rank cannot map its peers -> EVERY rank stays on the NCCL calls (the rank that did open switches the option off)"""
    res = _run_mailbox(bad_rank=1, backend_name="nccl", port_off=23)
    assert not any(use for _, use, _, _ in res)
    assert res[0][3] == [(b"p2p_scalars", 0)] and res[1][3] == []


def test_mailbox_negotiation_skipped_without_nccl():
    res = _run_mailbox(bad_rank=-1, backend_name="gloo", port_off=37)
    assert not any(use for _, use, _, _ in res) and all(op is None for _, _, op, _ in res)
