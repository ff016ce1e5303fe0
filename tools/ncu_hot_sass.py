#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an `ncu --page source --csv --print-source sass` dump.
    ncu -i X.ncu-rep --page source --csv --print-source sass > /tmp/src.csv ; python tools/ncu_hot_sass.py /tmp/src.csv [N]
The dump holds one table per captured launch; the LAST one is summarised."""
import csv
import sys


def main(path, top=22):
    rows = list(csv.reader(open(path, errors="ignore")))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    rows = rows[starts[-1]:]
    h = rows[0]
    src, smp, exe = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "(Not Issued)" not in c]
    body = [r for r in rows[1:] if len(r) > smp and r[smp].strip().isdigit()]
    tot = sum(int(r[smp]) for r in body) or 1
    print(f"# {len(body)} SASS instructions, {tot} warp-stall samples; top {top} by samples")
    print(f"# {'share':>6s} {'samples':>8s} {'executed':>9s}  instruction  [dominant stall reasons]")
    for r in sorted(body, key=lambda r: -int(r[smp]))[:top]:
        why = sorted(((int(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
        why = ", ".join(f"{n}:{c}" for c, n in why if c)
        print(f"  {100.0 * int(r[smp]) / tot:5.1f}% {r[smp]:>8s} {r[exe]:>9s}  {r[src][:70]:70s}  [{why}]")
    agg = {}
    for r in body:
        for i in stalls:
            agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i] or 0)
    print("# stall reasons over the kernel: " + ", ".join(f"{k} {100.0 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ops = {}
    for r in body:
        op = r[src].split()[0] if r[src].split() else ""
        if op.startswith("@"):
            op = r[src].split()[1]
        op = op.split(".")[0]
        if op in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UTCATOMSWS", "LDTM", "UTMAPF", "UTMACCTL", "RED", "ATOMG", "ATOM", "STG", "LDG"):
            ops[op] = ops.get(op, 0) + 1
    print("# mnemonic census (static): " + ", ".join(f"{k}×{v}" for k, v in sorted(ops.items())))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
