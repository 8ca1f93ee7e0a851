#!/usr/bin/env python
"""Per-kernel error table of the CUDA engine (or the oracle with --oracle) against a reference-executed golden.
Usage: python tools/refexec_report.py <solver> [suffix] [--oracle]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import refexec  # noqa: E402
from tests.refexec import Golden, OracleImpl, EngineImpl  # noqa: E402

solver = sys.argv[1]
suffix = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
g = Golden(solver, suffix)
impl = OracleImpl(g) if "--oracle" in sys.argv else EngineImpl(g, list_cap_liquid=256, list_cap_solid=256)
rows = {}


def check(idx, k, f, mine, gold):
    if f.startswith("hg_"):
        e, sc = (0.0 if np.array_equal(np.asarray(mine), gold) else 1.0), 1.0
        ab = e
    elif f in refexec.GLOB:
        ab, sc = abs(mine - gold), abs(gold)
    else:
        a, b = np.asarray(mine, np.float64).reshape(-1), np.asarray(gold, np.float64).reshape(-1)
        fin = np.isfinite(b)
        ab = float(np.max(np.abs(a[fin] - b[fin]))) if fin.any() and np.array_equal(np.isfinite(a), fin) else float("inf")
        sc = float(np.max(np.abs(b[fin]))) if fin.any() else 0.0
    key = (k, f)
    r = rows.setdefault(key, [0.0, 0.0, -1])
    if ab / max(sc, 1e-30) > r[0] / max(r[1], 1e-30) or r[2] < 0:
        rows[key] = [ab, sc, g.step_of(idx)]


refexec.replay(g, impl, check, stop=g.steps[-1]["event_end"])
print("%-28s %-16s %12s %12s %10s step" % ("kernel", "field", "max|err|", "max|ref|", "rel"))
for (k, f), (ab, sc, st) in sorted(rows.items(), key=lambda kv: -(kv[1][0] / max(kv[1][1], 1e-30))):
    print("%-28s %-16s %12.4e %12.4e %10.2e %d" % (k, f, ab, sc, ab / max(sc, 1e-30), st))
