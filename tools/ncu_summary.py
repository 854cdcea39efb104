#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: duration, DRAM bytes read + written, achieved bandwidth from the
ALGORITHMIC bytes (tools/prof_hbm_kernels.py writes them), fraction of the measured HBM peak, and the stall / pipe
counters the B200 profiling recipe names.

    ncu -i gpurun_out/r2_hbm.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv [gpurun_out/prof_hbm_alg_bytes.json]

The LAST profiled launch of each kernel name is reported (the first one is the cold launch)."""
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu.sum"]


def to_num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def scale(v, unit, kind):
    if v is None:
        return None
    if kind == "time":
        return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(unit, 1.0)
    if kind == "bytes":
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return v


def main():
    raw = sys.argv[1]
    alg = json.loads(Path(sys.argv[2]).read_text()) if len(sys.argv) > 2 else {}
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm = peaks.get("hbm_gbs", 6650.0)
    with open(raw) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    col = {h: i for i, h in enumerate(header)}
    rows = {}
    order = []
    for r in rd:
        if len(r) < len(header):
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        name = re.sub(r"^void ", "", name)
        if name not in rows:
            order.append(name)
        rows[name] = r
    print(f"# HBM peak used: {hbm} GB/s ({'MEASURED_PEAKS.json' if peaks else 'fallback'}); per-launch numbers of the last profiled launch")
    print(f"{'kernel':58} {'us':>8} {'dram rd MB':>10} {'dram wr MB':>10} {'alg MB':>8} {'alg GB/s':>9} {'of peak':>7} {'dram GB/s':>9} {'dram%':>6} {'L2hit%':>6} {'warps%':>6} {'regs':>4}")
    for name in order:
        r = rows[name]

        def get(m, kind=None):
            if m not in col:
                return None
            return scale(to_num(r[col[m]]), units[col[m]], kind)

        us = get("gpu__time_duration.sum", "time")
        rdb, wrb = get("dram__bytes_read.sum", "bytes"), get("dram__bytes_write.sum", "bytes")
        short = name.split("::")[-1]
        base = re.sub(r"<.*", "", short)
        a = alg.get(base)
        if a is None and base.startswith("ntx_"):
            a = None
        ags = (a / (us * 1e-6) / 1e9) if (a and us) else None
        dgs = ((rdb or 0) + (wrb or 0)) / (us * 1e-6) / 1e9 if us else None
        f = lambda v, p=1: "-" if v is None else f"{v:.{p}f}"
        print(f"{short[:58]:58} {f(us):>8} {f(rdb / 1e6 if rdb is not None else None):>10} {f(wrb / 1e6 if wrb is not None else None):>10} "
              f"{f(a / 1e6 if a else None):>8} {f(ags, 0):>9} {f(ags / hbm if ags else None, 3):>7} {f(dgs, 0):>9} "
              f"{f(get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):>6} {f(get('lts__t_sector_hit_rate.pct')):>6} "
              f"{f(get('sm__warps_active.avg.pct_of_peak_sustained_active')):>6} {f(get('launch__registers_per_thread'), 0):>4}")


if __name__ == "__main__":
    main()
