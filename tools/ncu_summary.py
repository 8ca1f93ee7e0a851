#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` output into the markdown table kept under profiles/ and (optionally) the
per-kernel dram-traffic json bench.py reads.  Usage: python tools/ncu_summary.py raw.csv [--traffic-json out.json --config c2]"""
import csv
import json
import sys

COLS = [("time us", "gpu__time_duration.sum", 1.0), ("DRAM read MB", "dram__bytes_read.sum", 1.0), ("DRAM write MB", "dram__bytes_write.sum", 1.0),
        ("regs", "launch__registers_per_thread", 1.0), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
        ("issue active %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0), ("L1 hit %", "l1tex__t_sector_hit_rate.pct", 1.0),
        ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0), ("L1 tp %", "l1tex__throughput.avg.pct_of_peak_sustained_active", 1.0),
        ("DRAM tp %", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0), ("warp-instr M", "smsp__inst_executed.sum", 1e-6)]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
         "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = ["| kernel | " + " | ".join(c[0] for c in COLS) + " | top stalls |", "|---|" + "---|" * (len(COLS) + 1)]
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        cells = []
        for _, key, mul in COLS:
            v = num(d.get(key, "")) if key in d else next((num(x) for k2, x in d.items() if k2.endswith(key)), None)
            key = key if key in d else next((k2 for k2 in d if k2.endswith(key)), key)
            if v is None:
                cells.append("-")
                continue
            v *= SCALE.get(u.get(key, ""), 1.0) * mul
            cells.append("%.1f" % v if v < 1000 else "%.0f" % v)
        st = sorted(((num(v) or 0.0, k.split("issue_stalled_")[1].split("_per")[0]) for k, v in d.items()
                     if "issue_stalled" in k and (k.endswith("per_warp_active.pct") or k.endswith("per_issue_active.ratio"))), reverse=True)[:3]
        out.append("| `%s` | %s | %s |" % (name, " | ".join(cells), ", ".join("%s %.1f" % (k, v) for v, k in st)))
        rd, wr = num(d.get("dram__bytes_read.sum", "")), num(d.get("dram__bytes_write.sum", ""))
        if rd is not None and wr is not None:
            t = (rd * SCALE.get(u["dram__bytes_read.sum"], 1.0) + wr * SCALE.get(u["dram__bytes_write.sum"], 1.0)) * 1e6
            wi = num(d.get("smsp__inst_executed.sum", ""))
            traffic.setdefault(name, {"dram_bytes": t, "warp_inst": wi, "time_us": num(d.get("gpu__time_duration.sum", "")) and
                                      num(d["gpu__time_duration.sum"]) * SCALE.get(u.get("gpu__time_duration.sum", ""), 1.0)})
    print("\n".join(out))
    if "--traffic-json" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic-json") + 1]
        cfg = sys.argv[sys.argv.index("--config") + 1] if "--config" in sys.argv else "c2"
        try:
            old = json.load(open(path))
        except (OSError, ValueError):
            old = {}
        old[cfg] = traffic
        json.dump(old, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
