#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by kernel.
    python tools/summarize_launches.py profiles/ncu_launches_r01.csv > profiles/ncu_launches_r01_summary.txt
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        unit = r[mu]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        name = r[kn].split("(")[0].replace("void ", "").replace("seg::", "")
        agg[name][r[mn]] += v * scale
        if r[mn] == "gpu__time_duration.sum":
            cnt[name] += 1
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print(f"# {path}: {sum(cnt.values())} launches, total device time {tot / 1e3:.2f} ms (cold-cache, serialised: compare SHARES)")
    print(f"# {'share':>6} {'time_us':>10} {'launches':>8} {'avg_us':>8} {'dram_MB/launch':>14} {'dram_GB/s':>9}  kernel")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t = a["gpu__time_duration.sum"]
        dram = (a.get("dram__bytes_read.sum", 0.0) + a.get("dram__bytes_write.sum", 0.0)) / max(cnt[name], 1) / 1e6
        gbs = dram * 1e6 * max(cnt[name], 1) / (t * 1e-6) / 1e9 if t > 0 else 0.0  # DRAM bytes / device time of the kernel's launches
        print(f"  {100 * t / tot:5.1f}% {t:10.1f} {cnt[name]:8d} {t / max(cnt[name], 1):8.1f} {dram:14.2f} {gbs:9.0f}  {name}")


if __name__ == "__main__":
    main(sys.argv[1])
