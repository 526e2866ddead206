#!/usr/bin/env python
"""Reduce an `ncu --csv --page raw` log to the handful of columns the rooflines are argued from.

    python tools/ncu_extract.py gpurun_out/x_raw.csv [--group] > profiles/x_summary.txt

One line per captured launch (or per kernel name with --group: launches, total time, summed DRAM bytes, time-weighted
percentages).  The per-launch numbers are cold-cache and serialised (profiler replay): compare shares, not absolutes."""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "us", "us"), ("dram__bytes_read.sum", "rd_MB", "MB"), ("dram__bytes_write.sum", "wr_MB", "MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", None),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", None),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%", None),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", None),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", None),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", None),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%", None),
        ("launch__registers_per_thread", "regs", None)]
UNIT = {"us": {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}, "MB": {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}}


def rows(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    if "Metric Name" in header:            # long format (`--csv` without `--page raw`): one row per launch and metric
        kn, mn, mu, mv = (header.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
        grid = header.index("Grid Size")
        launches = {}
        for r in rd:
            if len(r) <= mv:
                continue
            d = launches.setdefault(r[0], {"kernel": re.sub(r"^void |ptk::|\(.*$", "", r[kn]), "grid": r[grid]})
            for want, label, kind in COLS:
                if r[mn] == want:
                    d[label] = float(r[mv].replace(",", "")) * (UNIT[kind].get(r[mu], 1.0) if kind else 1.0)
        yield from launches.values()
        return
    units = next(rd)
    idx = {}
    for want, label, _ in COLS:
        if want in header:
            idx[label] = header.index(want)
            continue
        for i, h in enumerate(header):
            if h.endswith("." + want):
                idx[label] = i
                break
    kn = header.index("Kernel Name")
    grid, block = header.index("Grid Size"), header.index("Block Size")
    for r in rd:
        if len(r) <= kn:
            continue
        out = {"kernel": re.sub(r"^void |ptk::|\(.*$", "", r[kn]), "grid": r[grid], "block": r[block]}
        for want, label, kind in COLS:
            if label not in idx:
                continue
            try:
                v = float(r[idx[label]].replace(",", ""))
            except ValueError:
                continue
            if kind:
                v *= UNIT[kind].get(units[idx[label]], 1.0)
            out[label] = v
        yield out


def main():
    path = sys.argv[1]
    group = "--group" in sys.argv
    labels = [c[1] for c in COLS]
    data = list(rows(path))
    if not group:
        print("# kernel | grid | " + " | ".join(labels))
        for d in data:
            print("%-60s %-14s " % (d["kernel"][:60], d["grid"]) + " ".join("%s=%.4g" % (k, d[k]) for k in labels if k in d))
        return
    agg = {}
    for d in data:
        a = agg.setdefault(d["kernel"], {"n": 0})
        a["n"] += 1
        t = d.get("us", 0.0)
        for k in labels:
            if k not in d:
                continue
            if k in ("us", "rd_MB", "wr_MB"):
                a[k] = a.get(k, 0.0) + d[k]
            elif k == "regs":
                a[k] = d[k]
            else:
                a[k] = a.get(k, 0.0) + d[k] * t
    print("# kernel | launches | total us | DRAM read MB | DRAM write MB | achieved DRAM GB/s | time-weighted: " + " ".join(labels[3:-1]))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1].get("us", 0)):
        t = a.get("us", 1e-9)
        gbs = (a.get("rd_MB", 0) + a.get("wr_MB", 0)) / t * 1e3 if t > 0 else 0
        print("%-56s n=%-3d us=%-9.1f rd=%-8.1f wr=%-8.1f GB/s=%-7.0f " % (k[:56], a["n"], t, a.get("rd_MB", 0), a.get("wr_MB", 0), gbs) +
              " ".join("%s=%.1f" % (lb, a[lb] / t) for lb in labels[3:-1] if lb in a) + " regs=%d" % a.get("regs", 0))


if __name__ == "__main__":
    main()
