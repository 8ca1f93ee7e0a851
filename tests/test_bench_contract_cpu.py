"""bench.py's contract, the parts that run without a GPU: the reference arm prints exactly ONE JSON line on stdout with the keys the
driver reads (whatever the libraries write in between goes to stderr), non-zero ranks of a torchrun launch stay silent, and the
N > 1 launch re-executes itself once with NCCL's init log switched on."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.pop("NCCL_DEBUG", None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--config", "c2_small", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "liquid particle-steps/s, DFSPH dam-break" and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "dam_break" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--config", "c2_small", "--steps", "1", "--warmup", "0"],
             {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_multi_rank_launch_turns_nccl_init_log_on(monkeypatch):
    """a banner-level (or absent) NCCL_DEBUG is raised to INFO / INIT by ONE re-exec; an exported INFO is left alone"""
    sys.path.insert(0, ROOT)
    import bench
    calls = []

    def fake_execv(exe, argv):
        calls.append((exe, list(argv), os.environ.get("NCCL_DEBUG"), os.environ.get("NCCL_DEBUG_SUBSYS")))
        raise SystemExit(0)

    class Stop(Exception):
        pass

    def stop():
        raise Stop()
    monkeypatch.setattr(os, "execv", fake_execv)
    monkeypatch.setattr(bench, "claim_stdout", stop)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "2"])
    for preset, expect_exec in ((None, True), ("VERSION", True), ("WARN", True), ("INFO", False), ("TRACE", False)):
        calls.clear()
        for k in ("NCCL_DEBUG", "NCCL_DEBUG_SUBSYS", "WCSPH_BENCH_REEXEC"):
            monkeypatch.delenv(k, raising=False)
        monkeypatch.setenv("WORLD_SIZE", "2")
        if preset:
            monkeypatch.setenv("NCCL_DEBUG", preset)
        try:
            bench.main()
        except SystemExit:
            pass
        except Stop:
            pass
        assert bool(calls) == expect_exec, (preset, calls)
        if expect_exec:
            assert calls[0][2:] == ("INFO", "INIT") and os.environ.get("WCSPH_BENCH_REEXEC") == "1"
            calls.clear()
            try:
                bench.main()                       # the re-executed process does not exec again
            except (SystemExit, Stop):
                pass
            assert not calls
