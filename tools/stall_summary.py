"""Summarise `ncu --page source --csv` (SASS view): total stall samples by reason, and the top instructions."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    si = hdr.index("# Samples")
    src = hdr.index("Source")
    reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si] or 0) for r in data)
    print("total samples", tot)
    agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in reasons}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:28s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
    print("top instructions:")
    order = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:top]
    for i in sorted(order):
        r = data[i]
        why = sorted(((int(r[j] or 0), hdr[j]) for j in reasons), reverse=True)[:2]
        print(f"  #{i:5d} {int(r[si]):7d} {100 * int(r[si]) / max(tot, 1):5.1f}%  {r[src].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
