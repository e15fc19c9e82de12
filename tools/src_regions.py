#!/usr/bin/env python
"""Region summary of an `ncu --page source --csv` dump: consecutive SASS instructions with the same executed
count are merged; prints each region's share of executed warp-instructions and of stall samples.
    python tools/src_regions.py src.csv [min_share_pct] [--list a b]"""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
    hdr = rows[hi]
    body = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            break
        body.append(r)
    return hdr, body


def main():
    hdr, body = load(sys.argv[1])
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    if "--list" in sys.argv:
        k = sys.argv.index("--list")
        a, b = int(sys.argv[k + 1]), int(sys.argv[k + 2])
        for i in range(a, b + 1):
            print("%4d %.3e %6s  %s" % (i, int(body[i][ia]), body[i][isamp], body[i][isrc].strip()))
        return
    mn = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    tot = sum(int(r[ia]) for r in body)
    tots = sum(int(r[isamp]) for r in body) or 1
    print("total warp-instructions %.4e, samples %d, SASS length %d" % (tot, tots, len(body)))
    regions = []
    for i, r in enumerate(body):
        n = int(r[ia])
        if regions and regions[-1][2] == n:
            regions[-1][1] = i
            regions[-1][3] += int(r[isamp])
        else:
            regions.append([i, i, n, int(r[isamp])])
    for a, b, n, sm in regions:
        sh = 100.0 * n * (b - a + 1) / tot
        if sh > mn:
            print("%4d-%4d len %3d exec/inst %.3e share %5.2f%% samples %5.2f%%  first: %s" % (
                a, b, b - a + 1, n, sh, 100.0 * sm / tots, body[a][isrc].strip()[:60]))


if __name__ == "__main__":
    main()
