"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total device time and share."""
import collections
import csv
import re
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = re.sub(r"\(.*", "", r[ki])
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"{'kernel':72s} {'count':>6s} {'total us':>10s} {'share':>6s} {'us/launch':>10s}")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{n[:72]:72s} {c:6d} {v / 1e3:10.1f} {100 * v / tot:5.1f}% {v / 1e3 / c:10.2f}")
    print(f"{'TOTAL':72s} {sum(c for c, _ in agg.values()):6d} {tot / 1e3:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
