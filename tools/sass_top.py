"""Top stalled SASS instructions of an `ncu --page source --csv` export (tools/ncu_source.sh).
usage: python tools/sass_top.py <sass.csv> [n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    print(rows[0][:2])
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    idx = {h: hdr.index(h) for h in stalls}
    tot = sum(int(r[isamp]) for r in data)
    agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stalls}
    print("total samples", tot, "instructions", len(data))
    print("stall mix:", ", ".join(f"{h[6:]} {v / tot * 100:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
    for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp]))[:n]:
        top = max(stalls, key=lambda h: int(r[idx[h]] or 0))
        print(f"{i:5d} {int(r[isamp]):6d} {int(r[isamp]) / tot * 100:5.1f}% x{r[iex]:>9s} {top[6:]:14s} {r[isrc].strip()[:80]}")


if __name__ == "__main__":
    main()
