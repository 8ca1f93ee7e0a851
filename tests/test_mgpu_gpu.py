"""Multi-GPU parity: runs tests/mgpu_check.py under torchrun on 2 (and 4) GPUs when the box has them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_z_slab_ranks_match_oracle(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29530 + world), os.path.join(ROOT, "tests", "mgpu_check.py"),
           "12", "12", str(16 * world + 16), "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("solver,steps", [("sesph", 10), ("iisph", 6), ("pcisph", 6)])
def test_other_solvers_on_z_slab_ranks(solver, steps):
    """SESPH / IISPH / PCISPH over two slabs: each sweep is preceded by the halo of exactly the fields it gathers from j."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_check.py"),
           "10", "10", "32", str(steps), solver]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_z_slab_ranks_scene_in_motion(world):
    """slab ranks vs the oracle with the liquid MOVING through the slab faces (+z drift, stiff viscosity): migrants > 0 on every
    interior face, the divergence / viscosity / pressure loops take the oracle's iteration counts, rho / x within 1e-4."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29550 + world), os.path.join(ROOT, "tests", "mgpu_check.py"),
           "moving", "12", "12", "8", "12"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_checkpoint_restart_on_slab_ranks():
    """save / load_state with world_size 2: the restore re-partitions the liquids by their restored cell layers."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29561", os.path.join(ROOT, "tests", "mgpu_check.py"), "checkpoint"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_surface_reconstruction_on_slab_ranks():
    """MCGrid.update_grid / cal_surface_point / marching_cube with world_size 2: per-rank shares of the colour field, summed."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "mgpu_check.py"), "surface"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
