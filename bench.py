#!/usr/bin/env python
"""bench.py -- liquid particle-steps/s of the DFSPH dam-break hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c5|...]

One "step" = one pass of the reference main-loop body dfsph.py:606-617 (grid build ... update_pos)
over the whole scene.  At N=1 the workload is BASELINE configs[1]: DFSPH dam-break, 1M liquid
particles, fp32 (`scenes.dam_break(100,100,100)`).  Under torchrun (N>1) every rank runs one
process on its own GPU.

`value`   : whole-job particle-steps/s, state resident in HBM, CUDA-event timed, max over ranks.
`e2e`     : the same K steps (fresh scene + the same warm-up) driven through the reference-facing Field API with HOST state --
            pos/vel H2D from pinned memory before, pos/vel D2H after, every step, in the region.
stdout    : exactly one JSON line; everything else (NCCL's init log included, switched on for N > 1) goes to stderr.
`roofline`: dominant kernel, algorithmic bytes / CUDA-event duration measured in a
            separate profiled pass of the same K steps in this process.
`cpu_baseline` / `--impl reference`: the CPU oracle (a port: the reference is Taichi DSL and
            Taichi is not installable here) on all host threads, bounded sample of the same scene.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (solver, (nx, ny, nz), description)
    "c2": ("dfsph", (100, 100, 100), "DFSPH dam-break 1M particles fp32 (BASELINE configs[1])"),
    "c5": ("dfsph", (200, 200, 400), "DFSPH dam-break 16M particles fp32 (BASELINE configs[4])"),
    "c5_rank": ("dfsph", (200, 200, 50), "DFSPH dam-break 2M particles: the share of ONE rank of configs[4] on 8 z-slabs, on one GPU (what a slab rank could reach with no exchange at all)"),
    "c5_2of8": ("dfsph", (200, 200, 100), "DFSPH dam-break 4M particles: two of the eight slabs of configs[4] (--gpus 2 reproduces the per-rank load of the 8-GPU run)"),
    "c2_small": ("dfsph", (40, 40, 40), "DFSPH dam-break 64k particles (bounded CPU sample of configs[1])"),
    # parity-test configurations of BASELINE.json, runnable here for the record (not the headline line)
    "c3": ("pcisph", (200, 100, 200), "PCISPH 4M particles + Akinci surface tension (BASELINE configs[2])"),
    "c4": ("iisph", (200, 100, 100), "IISPH 2M particles + Weiler implicit-viscosity PCG (BASELINE configs[3])"),
    "c1": ("sesph", None, "SESPH 3D dam-break ~8k particles, as shipped (BASELINE configs[0])"),
    # the north star's "density+force pass" in isolation: SESPH is exactly density(+EOS) sweep + force sweep + integrate
    "c1_1m": ("sesph", (100, 100, 100), "SESPH dam-break 1M particles fp32: density+force pass at the size of configs[1]"),
}
CPU_SAMPLE_SMALL = (40, 40, 40)
FP32_PEAK_TFLOPS = 74.0        # 148 SMs x 128 lanes x 2 flop x 1.965 GHz, non-tensor FP32
# in-range pair work of the sweeps, SURVEY 8(d): flop per in-range pair (density 20, gradW-based sums 35-45, viscosity Ax 60)
PAIR_FLOPS = {"k_dfsph_drho": 40, "k_dfsph_velcorrect": 40, "k_dfsph_head": 75, "k_visc_Ad": 60, "k_visc_minv_residual": 120,
              "k_vorticity_fused": 110, "k_sesph_density": 20, "k_sesph_force": 60}


def cpu_sample_dims(dims):
    """the CPU arm runs the SAME scene as the GPU arm when the reference's data structures fit the host (8 KiB of neighbour table
    + 256 B of bucket slots per liquid particle: 9 GB at 1M); otherwise the largest dam-break that does."""
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 0
    nl = dims[0] * dims[1] * dims[2]
    need = nl * (2048 * 4 + 64 * 4 * 1.2 + 400)
    if need < 0.6 * avail and nl <= 1000000:
        return dims, True
    if 1000000 * (2048 * 4 + 64 * 4 * 1.2 + 400) < 0.6 * avail:
        return (100, 100, 100), dims == (100, 100, 100)
    return CPU_SAMPLE_SMALL, False


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)       # 50 steps after 10 warm-up
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-slab-parity", action="store_true", help="N > 1: skip the moving-scene slab-vs-oracle check that precedes the timed region")
    ap.add_argument("--quick", action="store_true", help="N = 1: skip the developed-flow figure and the 16M strong-scaling base")
    ap.add_argument("--no-graph", action="store_true",
                    help="stream-ordered launches with host-driven loops instead of one CUDA graph per step (same kernels): "
                         "for ncu launch lists -- ncu cannot see kernel nodes of graphs that hold conditional nodes")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  An in-process NVML thread
    (pynvml, ~2 ms period: a 30 ms timed region still gets samples); `nvidia-smi -lms` as the fallback when NVML will not load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.nv = self.h = self.thread = self.p = self.f = None
        self.sm, self.reason_bits, self.run = [], 0, False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:                                               # CUDA_VISIBLE_DEVICES may renumber: go by UUID
                uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            import threading
            self.run = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.run = False
            self.thread.join(timeout=2)
            nv = self.nv
            names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
            out["reasons"] = sorted(n for n, a in names if self.reason_bits & int(getattr(nv, a, 0)))
            if self.sm:
                out["sm_mhz"], out["samples"] = float(np.median(self.sm)), len(self.sm)
            out["sm_max_mhz"] = self.sm_max
            out["source"] = "nvml thread, 2 ms period, inside the timed region"
            return out
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        out["source"] = "nvidia-smi -lms 100"
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(N, NL, ncells):
    """compulsory bytes per launch (every input array read once, every output
    written once, vec3 = 12 B, no index / neighbour-list traffic), per kernel of the DFSPH step."""
    return {
        "k_dfsph_head": 12 * N + 12 * NL + 12 * NL,                   # density + alpha + warm-start Drho/Dt: pos, vel -> rho, alpha, adv_rho
        "k_visc_minv_residual": 12 * N + 4 * NL + 24 * NL + 36 * NL + 24 * NL,
        "k_vorticity_fused": 12 * N + 4 * NL + 24 * NL + 36 * NL + 24 * NL + 4 * NL,
        "(k_dfsph_density_alpha<true, true>)": 12 * N + 8 * NL,       # compute_density + compute_dfsph_coff fused: pos -> rho, alpha
        "(k_dfsph_drho<0, true, false, false>)": 12 * N + 16 * NL,    # update_drho_divergence: pos, vel -> adv_rho
        "(k_dfsph_drho<0, false, true, false>)": 12 * N + 16 * NL,
        "(k_dfsph_drho<0, false, false, true>)": 12 * N + 16 * NL,
        "(k_dfsph_drho<1, false, true, false>)": 12 * N + 20 * NL,    # update_drho_pressure: + rho
        "(k_dfsph_drho<1, false, false, true>)": 12 * N + 20 * NL,
        "k_dfsph_velcorrect<0>": 12 * N + 40 * NL,                    # pos, (alpha, adv_rho), vel rmw, kappa rmw
        "k_dfsph_velcorrect<1>": 12 * N + 40 * NL,
        # k_dfsph_velcorrect<2> (warmstart_pressure) is the dead Q13 branch: every thread returns after one compare -> no figure
        "k_dfsph_velcorrect<3>": 12 * N + 40 * NL,
        "(k_dfsph_velcorrect<1, true>)": 12 * N + 40 * NL,            # the same sweep with kfac_j packed in pos.w (one GPU)
        "(k_dfsph_velcorrect<3, true>)": 12 * N + 40 * NL,
        "k_visc_minv": 12 * N + 4 * NL + 36 * NL,
        "k_visc_residual": 12 * N + 4 * NL + 12 * NL + 12 * NL + 24 * NL,
        "k_visc_Ad": 12 * N + 4 * NL + 12 * NL + 12 * NL,            # get_viscosity_Ax
        "k_visc_update": (36 + 5 * 12) * NL // 2 + 36 * NL,
        "k_vorticity": 12 * N + 4 * NL + 24 * NL + 24 * NL,
        "k_sesph_density<true>": 12 * N + 8 * NL,                    # density + EOS
        "k_sesph_density<false>": 12 * N + 4 * NL,
        "k_sesph_force": 12 * N + 32 * NL,                           # pos, (vel, rho, p) -> d_vel
        "k_build_lists": 12 * N + 4 * NL,                            # neighbour query: pos -> neighborCount (lists are not compulsory traffic)
        "k_permute": 2 * 60 * NL,                                    # reorder of the persistent state (pos, vel, omega, vel_guess, kappa, kappa_v, pressure, id)
        "cub_radix_sort": 16 * NL,
        "k_keys": 12 * NL + 4 * NL,
    }


def roofline_from_rows(rows, ab, ncu, peak, peak_src, K):
    """rows: {profiler name: (launches, total ms)} of K steps; ab: algorithmic bytes per launch by profiler name; ncu:
    dram bytes per launch by ncu kernel name.  The dominant kernel is the kernel FAMILY (all template instantiations of
    one __global__ function: they walk the same pairs and differ in one fused term) with the largest share of the step."""
    def family(nm):
        return nm.strip("()").split("<")[0]

    def ncu_name(nm):
        return nm.strip("()").replace("false", "0").replace("true", "1")
    # the profiler label "k_build_lists" covers whichever list-build instantiation ran (k_build_lists2<SLAB, S> since round 2)
    ncu = dict(ncu)
    for k in list(ncu):
        if k.startswith("k_build_lists"):
            ncu.setdefault("k_build_lists", ncu[k])
    tot = sum(v[1] for v in rows.values())
    fam = {}
    for name, (n, kms) in rows.items():
        f = fam.setdefault(family(name), {"ms": 0.0, "launches": 0, "bytes": 0.0, "traffic": 0.0, "traffic_launches": 0, "has_bytes": True})
        f["ms"] += kms
        f["launches"] += n
        if ab.get(name):
            f["bytes"] += ab[name] * n
        else:
            f["has_bytes"] = False
        if ncu_name(name) in ncu:
            f["traffic"] += ncu[ncu_name(name)] * n
            f["traffic_launches"] += n

    def roof_of(fname):
        f = fam[fname]
        ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9 if f["has_bytes"] and f["bytes"] else None
        return {"kernel": fname, "achieved": ach, "frac": (ach / peak) if ach else None,
                "traffic": (f["traffic"] / f["traffic_launches"]) if f["traffic_launches"] else None,
                "algorithmic_bytes_per_launch": (f["bytes"] / f["launches"]) if f["has_bytes"] else None,
                "avg_launch_ms": f["ms"] / f["launches"], "launches_per_step": f["launches"] / K, "share_of_step": f["ms"] / tot}
    ranked = sorted((k for k, f in fam.items() if f["has_bytes"] and f["bytes"]), key=lambda k: -fam[k]["ms"]) or sorted(fam, key=lambda k: -fam[k]["ms"])
    roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src}
    roof.update(roof_of(ranked[0]))
    roof["next"] = [roof_of(k) for k in ranked[1:4]]              # the following three families, same accounting
    roof["note"] = ("neighbour sweeps are FP32-issue / L1-latency bound, not HBM-bound (SURVEY fact 10, profiles/r01_ncu_full_sweeps_1M.md); "
                    "frac is against the HBM roof as the metric demands; traffic = ncu dram bytes per launch, launch-weighted over the family")
    return roof


def build_engine(solver, dims, world=1, rank=0):
    from wcsph_b200 import scenes
    import importlib
    mod = importlib.import_module("wcsph_b200." + solver)
    pts, nl = scenes.dam_break(*dims) if dims else getattr(scenes, "scene_" + solver)()
    if world > 1:
        mod.init_scene(pts, nl, world_size=world, rank=rank)     # z-slab rank: NCCL halo exchange inside the library
    else:
        mod.init_scene(pts, nl)
    mod.reset_param()
    if solver == "pcisph" and dims:
        mod.set_tension(0.1, 0.05)               # configs[2]: gamma = 0.1
    return mod, pts, nl


def profile_report(pd):
    from wcsph_b200 import _lib
    buf = C.create_string_buffer(1 << 16)
    _lib.check(_lib.load().wcsph_profile_report(pd._ctx, buf, len(buf)))
    rows = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split("\t")
        rows[name] = (int(n), float(ms))
    return rows


def cpu_baseline(steps=3, warmup=1, threads=None, dims=(100, 100, 100)):
    """oracle (kind "port") on all host threads, bounded sample: a few steps of the SAME dam-break scene when it fits the host"""
    from oracle.oracle import Oracle
    from wcsph_b200 import scenes
    threads = threads or os.cpu_count() or 1
    sample, same = cpu_sample_dims(dims)
    pts, nl = scenes.dam_break(*sample)
    o = Oracle("dfsph", pts, nl, threads=threads)
    for _ in range(warmup):
        o.step()
    t = time.perf_counter()
    its = []
    for _ in range(steps):
        o.step()
        its.append((o.flag("vs_iter"), o.flag("dv_iter"), o.flag("pr_iter")))
    dt = time.perf_counter() - t
    return {"value": nl * steps / dt, "unit": "particle-steps/s", "cores": threads, "kind": "port", "same_scene_as_gpu_arm": bool(same),
            "sample": "dam_break%s = %d liquid + %d boundary particles, %d steps after %d warm-up, reference data structures "
                      "(64-slot buckets, 2048-wide neighbour table), OpenMP over particles; iters(vs,dv,pr)=%s"
                      % (str(sample), nl, len(pts) - nl, steps, warmup, str(its[-1])),
            "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle port; Taichi is unavailable)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config or "c2"            # the CPU arm times configs[1]'s scene at every N (16M needs 131 GB of neighbour table)
    cb = cpu_baseline(steps=args.steps, warmup=args.warmup, dims=CONFIGS[cfg][1] or (100, 100, 100))
    line = {"impl": "reference", "metric": "liquid particle-steps/s, DFSPH dam-break", "value": cb["value"], "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[cfg][2], "solver": "dfsph",
                       "sample": "each step is one DFSPH step of the bounded sample " + cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def pair_counts(pd):
    from wcsph_b200 import _lib
    out = (C.c_longlong * 4)()
    _lib.check(_lib.load().wcsph_pair_counts(pd._ctx, C.byref(out)))
    return int(out[0]), int(out[1]), int(out[2]), int(out[3])


def ncu_table(cfg):
    """per-kernel ncu figures committed under profiles/ (dram bytes, warp instructions per launch); the capture command is in
    profiles/ncu_traffic.json's "_source".  Old files hold a bare number = dram bytes."""
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(tp):
        return {}, None
    j = json.load(open(tp))
    tab = {}
    for k, v in j.get(cfg, {}).items():
        tab[k] = v if isinstance(v, dict) else {"dram_bytes": v, "warp_inst": None}
    return tab, j.get("_source")


def timed_steps(fused, K, W, barrier, local, world, dist):
    """W warm-up + K timed fused steps -> (ms max over ranks, clocks)"""
    import torch
    fused(W)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    fused(K)                                   # dfsph: K CUDA-graph launches queued back to back, no host sync inside
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), clocks


def iter_stats(iters):
    if not iters:
        return {}
    a = np.asarray(iters, dtype=np.float64)
    return {"iters_mean_vs": float(a[:, 0].mean()), "iters_mean_dv": float(a[:, 1].mean()), "iters_mean_pr": float(a[:, 2].mean()),
            "iters_last_vs": int(a[-1, 0]), "iters_last_dv": int(a[-1, 1]), "iters_last_pr": int(a[-1, 2])}


_JSON_OUT = [None]


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep the real stdout for it and point fd 1 at stderr for everything else (NCCL prints
    its version banner to stdout whatever NCCL_DEBUG_FILE says; libraries and warnings do the same now and then)."""
    if _JSON_OUT[0] is None:
        sys.stdout.flush()
        _JSON_OUT[0] = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT[0] or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    if (int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.impl != "reference"
            and os.environ.get("NCCL_DEBUG", "WARN").upper() in ("VERSION", "WARN", "NONE", "") and os.environ.get("WCSPH_BENCH_REEXEC") != "1"):
        # NCCL's own log (communicator size, transport) belongs in stderr of every N > 1 run.  The workers torchrun starts on this
        # image arrive with NCCL_DEBUG at banner level (only "NCCL version ..." appears, and a setdefault() does nothing -- measured,
        # tools/proto/nccl_log_probe.py): raise it to INFO / INIT and re-exec this rank once so that NCCL sees it from its first
        # getenv.  An INFO / TRACE level the caller exported is left alone.
        os.environ.update(NCCL_DEBUG="INFO", NCCL_DEBUG_SUBSYS="INIT", WCSPH_BENCH_REEXEC="1")
        sys.stdout.flush(); sys.stderr.flush()
        os.execv(sys.executable, [sys.executable] + sys.argv)
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        # (NCCL_DEBUG: see the top of main(); NCCL logs to fd 1, which claim_stdout() has pointed at stderr)
        # (no NCCL_DEBUG_FILE: NCCL logs to fd 1, which claim_stdout() has pointed at stderr -- opening /dev/stderr as a FILE would
        # truncate a redirected log and write over it from offset 0)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- z-slab parity on a scene in motion, OUTSIDE any timed region (N > 1): every SCALE line carries it --------------
    slab_parity = None
    if world > 1 and not args.no_slab_parity:
        from tests.mgpu_check import slab_parity as _slab_parity
        slab_parity = _slab_parity(world, rank)
        barrier()

    # N = 1: BASELINE configs[1] (1M).  N > 1: BASELINE configs[4] (16M), one z-slab per GPU, strong scaling
    cfg = args.config or ("c2" if world == 1 else "c5")
    solver, dims, desc = CONFIGS[cfg]
    def fresh_engine():
        m_, pts_, nl_ = build_engine(solver, dims, world, rank)
        pd_ = m_.particle_data
        if os.environ.get("WCSPH_HALO_OVERLAP") is not None:          # A/B switch for tools/ runs
            from wcsph_b200 import _lib as _l
            _l.check(_l.load().wcsph_set_option(pd_._ctx, b"halo_overlap", int(os.environ["WCSPH_HALO_OVERLAP"])))
        if args.no_graph and solver == "dfsph":
            m_.set_graph(False)
        return m_, pts_, nl_, pd_

    mod, pts, nl, pd = fresh_engine()
    N = len(pts)
    K, W = args.steps, max(args.warmup, 3)

    # ---- resident pass (value) ---------------------------------------------------------
    is_dfsph = solver == "dfsph"
    fused = (lambda n: mod.step_fused(n, fetch_iters=False)) if is_dfsph else (lambda n: mod.step_fused(n))
    fused(W)
    barrier()
    pd.launch_count(reset=True)
    ms_max, clocks = timed_steps(fused, K, 0, barrier, local, world, dist)
    iters = mod.iters_log(K) if is_dfsph else [(getattr(mod, "vs_iter", 0), getattr(mod, "dv_iter", 0), getattr(mod, "pr_iter", 0))]
    launches = pd.launch_count()
    pd.check()                                 # raises if the device dropped pairs during the timed steps
    flags = pd.hash_grid.status()
    value = nl * K / (ms_max * 1e-3)          # nl is the WHOLE scene's liquid count: all ranks step it together
    from wcsph_b200 import _lib
    L = _lib.load()
    ctx = pd._ctx
    mig0 = (C.c_longlong * 5)()
    _lib.check(L.wcsph_migration_counts(ctx, C.byref(mig0)))

    # ---- e2e pass: host-resident pos / vel cross PCIe every step ----------------------
    # It times the SAME K steps of the flow as the resident pass: the scene is set up again and warmed up by the same W steps (a
    # dam break costs more per step as it develops -- steps 27..47 of the 1M scene take 15 % longer than steps 5..25 -- so an e2e
    # pass that simply carried on would mix the cost of the copies with the cost of a later flow).
    pd = None
    mod, pts, nl, pd = fresh_engine()
    ctx = pd._ctx
    fused(W)
    barrier()
    if world == 1:
        # reference-facing Field API (reference order): pos.from_numpy / vel.from_numpy, step, to_numpy
        pos_h = torch.empty((N, 3), dtype=torch.float32).pin_memory()
        vel_h = torch.empty((nl, 3), dtype=torch.float32).pin_memory()
        pos_h.copy_(torch.from_numpy(pd.pos.to_numpy()))
        vel_h.copy_(torch.from_numpy(pd.vel.to_numpy()))

        def e2e_step():
            _lib.check(L.wcsph_field_set_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4))
            _lib.check(L.wcsph_field_set_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4))
            fused(1)
            _lib.check(L.wcsph_field_get_async(ctx, b"pos", pos_h.data_ptr(), pos_h.numel() * 4))
            _lib.check(L.wcsph_field_get_async(ctx, b"vel", vel_h.data_ptr(), vel_h.numel() * 4))
            pd.sync()          # the host owns the state again (the reference's pos.to_numpy(), dfsph.py:645)
            # bytes that cross PCIe: field_set copies the LIQUID rows of pos (solids are static, Q23) + vel; field_get returns all of pos + vel
            return (nl * 3 + nl * 3) * 4, (N * 3 + nl * 3) * 4
        e2e_api = "Field.from_numpy/to_numpy path (wcsph_field_set_async / _get_async, pinned, reference order) around dfsph.step_fused"
    else:
        # z-slab ranks: each rank's host keeps the state of ITS slab (cell-sorted device views of the
        # owned particles, wcsph_field_device): H2D before, D2H after, every step
        cap = pd.slab_plan["cap_own"]
        pos_h = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
        vel_h = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
        n0 = pd.pos.to_torch().shape[0]
        pos_h[:n0].copy_(pd.pos.to_torch()); vel_h[:n0].copy_(pd.vel.to_torch())
        torch.cuda.synchronize()
        state = {"n": n0}

        def e2e_step():
            n = state["n"]
            pd.pos.to_torch().copy_(pos_h[:n], non_blocking=True)
            pd.vel.to_torch().copy_(vel_h[:n], non_blocking=True)
            mod.step_fused(1, fetch_iters=False)
            pv, vv = pd.pos.to_torch(), pd.vel.to_torch()          # slab population changes with migration
            n2 = pv.shape[0]
            pos_h[:n2].copy_(pv, non_blocking=True); vel_h[:n2].copy_(vv, non_blocking=True)
            pd.sync()
            state["n"] = n2
            return 2 * n * 16, 2 * n2 * 16
        e2e_api = "per-rank slab state through wcsph_field_device views (pinned host <-> device, sorted order) around dfsph.step_fused"

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    h2d = d2h = 0
    for _ in range(K):
        a, b2 = e2e_step()
        h2d += a; d2h += b2
    ev1.record()
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([e2e_ms, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
    if world > 1:
        tm = t[:1].clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        tb = t[1:].clone(); dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        t = torch.cat([tm, tb])
    e2e_value = nl * K / (float(t[0].item()) * 1e-3)
    h2d = int(t[1].item() / K)
    d2h = int(t[2].item() / K)

    # ---- profiled pass: per-kernel CUDA-event durations (roofline) ------------------
    roof = None
    kernels = {}
    comm = None
    _lib.check(L.wcsph_profile(ctx, 1))          # every rank steps (the pass contains collectives); rank 0 reports
    for _ in range(K):
        fused(1)
    rows = profile_report(pd)
    _lib.check(L.wcsph_profile(ctx, 0))
    pl, ps, maxl, maxs = pair_counts(pd)
    if rank == 0:
        tot = sum(v[1] for v in rows.values())
        n_own = pd.pos.to_torch().shape[0]       # particles this rank sweeps (all liquids on one GPU)
        ab = algorithmic_bytes(n_own + (N - nl), n_own, int(np.prod(pd.hash_grid.blockSize[0])))
        peak, peak_src = measured_peak()
        for name, (n, kms) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
            avg = kms / n
            b = ab.get(name)
            kernels[name] = {"launches_per_step": n / K, "avg_ms": avg, "share": kms / tot,
                             "alg_GBps": (b / (avg * 1e-3) / 1e9) if b else None}
        ncu, ncu_src = ncu_table(cfg)
        roof = roofline_from_rows(rows, ab, {k: v["dram_bytes"] for k, v in ncu.items()}, peak, peak_src, K)
        roof["traffic_source"] = ncu_src or "profiles/ncu_traffic.json (ncu --set full capture of this configuration; null if none committed for it)"
        # second figure of SURVEY 8(d): FP32 work of the dominant sweep family.  pairs = in-range pairs of the compact lists
        # (measured: wcsph_pair_counts), flop per pair from SURVEY 8(d)'s table; peak 74 TFLOP/s = 148 SM x 128 lanes x 2 x 1.965 GHz
        fam = roof["kernel"]
        pf = PAIR_FLOPS.get(fam)
        if pf:
            pairs = pl + ps
            ach = pairs * pf / (roof["avg_launch_ms"] * 1e-3) / 1e12
            roof["fp32"] = {"pair_flops": pf, "pairs_per_launch": pairs, "achieved_tflops": ach, "peak_tflops": FP32_PEAK_TFLOPS,
                            "frac_of_74": ach / FP32_PEAK_TFLOPS}
        # third figure: issue slots.  warp instructions per launch from the committed ncu capture / (duration x 148 SMs x 4 schedulers x clock)
        wi = [v["warp_inst"] for k, v in ncu.items() if k.split("<")[0] == fam and v.get("warp_inst")]
        if wi:
            clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
            roof["issue_frac"] = float(np.mean(wi)) / (roof["avg_launch_ms"] * 1e-3 * 148 * 4 * clk)
        roof["list_stats"] = {"liquid_pairs": pl, "solid_pairs": ps, "pairs_per_particle": (pl + ps) / max(n_own, 1),
                              "longest_liquid_list": maxl, "longest_solid_list": maxs}
    if world > 1:
        mig1 = (C.c_longlong * 5)()
        _lib.check(L.wcsph_migration_counts(ctx, C.byref(mig1)))
        halo_ms = rows.get("nccl_halo", (0, 0.0))
        ci = (C.c_int * 4)()
        _lib.check(L.wcsph_comm_info(ctx, C.byref(ci)))
        comm = {"halo_exchanges_per_step": halo_ms[0] / K, "nccl_halo_ms_per_step": halo_ms[1] / K,
                "nccl_counts_ms_per_step": rows.get("nccl_counts(+host sync)", (0, 0.0))[1] / K,
                "nccl_migrate_ms_per_step": rows.get("nccl_migrate", (0, 0.0))[1] / K,
                "migrated_rank0_total": [int(mig1[k]) for k in range(4)],
                "nccl_nranks": int(ci[0]), "nccl_rank": int(ci[1]), "nccl_version": int(ci[2]),      # ncclCommCount / ncclCommUserRank / ncclGetVersion
                "peer_mailboxes": bool(ci[3]),
                "note": "rank 0's event-timed calls in the profiled pass; the halo of a sweep overlaps its interior launch and is followed "
                        "by the boundary strips on the same side stream; with peer_mailboxes the scalar all-reduces (inside k_finalize) and "
                        "the neighbour counts ('nccl_counts') are 8-byte peer stores over NVLink, not NCCL calls"}

    # ---- N = 1 extras: same-workload base of the strong-scaling curve, developed flow, CPU baseline ------------------
    strong_base = developed = None
    cb = None
    if world == 1 and rank == 0 and cfg == "c2" and not args.quick:
        # developed flow: the same engine 400 steps further into the collapse (the loops iterate more than from rest)
        fused(400)
        pd.sync()
        ms_d, clk_d = timed_steps(fused, K, 0, barrier, local, world, dist)
        it_d = mod.iters_log(K)
        developed = {"steps_in": W + 2 * K + 2 + 400, "ms_per_step": ms_d / K, "value": nl * K / (ms_d * 1e-3), "unit": "particle-steps/s",
                     "clocks_sm_mhz": clk_d.get("sm_mhz")}
        developed.update(iter_stats(it_d))
        pd.check()
        # BASELINE configs[4] (16M) on this one GPU: the N = 1 point of the N = 2/4/8 curve (the 1M engine stays resident: 1 + 15 GB)
        s5, d5, desc5 = CONFIGS["c5"]
        mod5, pts5, nl5 = build_engine(s5, d5, 1, 0)
        f5 = lambda n: mod5.step_fused(n, fetch_iters=False)
        ms5, clk5 = timed_steps(f5, K, W, barrier, local, world, dist)
        it5 = mod5.iters_log(K)
        mod5.particle_data.check()
        strong_base = {"workload": desc5, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms5 / K, "value": nl5 * K / (ms5 * 1e-3),
                       "unit": "particle-steps/s", "clocks_sm_mhz": clk5.get("sm_mhz"),
                       "note": "same scene / steps / warm-up as the --gpus 2/4/8 lines: divide their value by this one for the strong-scaling speed-up"}
        strong_base.update(iter_stats(it5))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(dims=dims or (100, 100, 100))

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    config = {"workload": desc, "solver": solver, "liquid_particles": nl, "boundary_particles": N - nl,
              "scene": "scenes.dam_break%s" % (dims,),
              "parallelism": "1 GPU, one CUDA graph per step" if world == 1 else
                             "%d z-slabs (one process per GPU), NCCL halo exchange per neighbour pass + per-step migration, fixed total scene" % world,
              "l2": "working set per GPU (state + neighbour lists, >= 0.7 GB) exceeds the 126 MB L2; no flush needed",
              "status_flags": flags, "start": "block at rest (t = 0) + %d warm-up steps" % W}
    config.update(iter_stats(iters))
    line = {
        "metric": "liquid particle-steps/s, %s dam-break" % solver.upper(), "value": value, "unit": "particle-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        # one label for the whole 1/2/4/8 curve: the scene is fixed per curve.  N = 1 carries BOTH the 1M headline (configs[1]) and,
        # under "strong_base", the 16M scene of the N > 1 lines on one GPU.
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": e2e_api},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "kernels": kernels,
        "cpu_baseline": cb,
    }
    if developed:
        line["developed_flow"] = developed
    if strong_base:
        line["strong_base"] = strong_base
    if slab_parity:
        line["slab_parity"] = slab_parity
    if comm:
        line["comm"] = comm
    emit(line)


if __name__ == "__main__":
    main()
