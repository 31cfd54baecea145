#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python srcprof.py src.csv [N]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


def main(path, top_n=45):
    rows = list(csv.reader(open(path)))
    cur = None
    hdr = None
    out = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] != "" and hdr and len(r) == len(hdr) and r[0].isdigit():
            out.append((cur, r))
    ii = hdr.index("Instructions Executed")
    ti = hdr.index("Thread Instructions Executed")
    si = hdr.index("# Samples")
    tot = sum(num(r[ii]) for _, r in out)
    tots = sum(num(r[si]) for _, r in out)
    tthr = sum(num(r[ti]) for _, r in out)
    print(f"total warp inst {tot / 1e9:.2f} G, thread inst {tthr / 1e9:.2f} G, avg active threads {tthr / max(tot, 1):.2f}, "
          f"samples {tots}")
    top = sorted(out, key=lambda fr: -num(fr[1][si]))[:top_n]
    for f, r in top:
        inst, th = num(r[ii]), num(r[ti])
        print(f"{f.split('/')[-1][:14]:14}:{r[0]:>4} inst={inst / 1e6:8.1f}M ({100 * inst / max(tot, 1):4.1f}%) "
              f"thr/inst={th / max(inst, 1):5.1f} samp={100 * num(r[si]) / max(tots, 1):5.1f}% | {r[1].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
