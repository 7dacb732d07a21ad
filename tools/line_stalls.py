"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line: samples, share, top stall reasons.
usage: python tools/line_stalls.py report_cs.csv [top]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path, errors="replace")))
    cur_file, hdr, out = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[2] != "-":          # SASS rows carry an address; CUDA-line rows have "-"
            continue
        si = hdr.index("# Samples")
        n = int(r[si] or 0)
        if n == 0:
            continue
        reasons = [(int(r[i] or 0), hdr[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        reasons.sort(reverse=True)
        ie = int(r[hdr.index("Instructions Executed")] or 0)
        out.append((n, cur_file, int(r[0]), r[1].strip()[:90], reasons[:3], ie))
    tot = sum(o[0] for o in out)
    print("total samples", tot)
    for n, f, ln, src, why, ie in sorted(out, reverse=True)[:top]:
        w = " ".join(f"{k[6:]}:{v}" for v, k in why if v)
        print(f"{n:6d} {100 * n / tot:5.1f}%  {f}:{ln:<5d} inst {ie:9d}  {src:90s} | {w}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
