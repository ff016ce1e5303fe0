#!/usr/bin/env python
"""ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) -> the per-launch DRAM
traffic of the tcgen05 conv kernels that bench.py reports as roofline.traffic.
    python tools/make_traffic_json.py gpurun_out/launches_r01c.csv profiles/roofline_traffic_r01.json
"""
import csv
import json
import sys


def main(src, dst):
    rows = list(csv.reader(open(src, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= mv or "conv_gemm_tc" not in r[kn] or r[mn] not in tot:
            continue
        try:
            tot[r[mn]] += float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
        except ValueError:
            continue
        n += r[mn] == "gpu__time_duration.sum"
    out = {"kernel": "conv_gemm_tc* (all tcgen05 conv launches of the traced steps)", "launches": n,
           "dram_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1),
           "dram_read_bytes_per_launch": tot["dram__bytes_read.sum"] / max(n, 1),
           "dram_write_bytes_per_launch": tot["dram__bytes_write.sum"] / max(n, 1),
           "avg_us_per_launch_under_ncu": tot["gpu__time_duration.sum"] / max(n, 1), "source": src}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
