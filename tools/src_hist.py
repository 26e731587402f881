#!/usr/bin/env python
"""Executed warp instructions per CUDA source line, from
   ncu -i X.ncu-rep --page source --csv --kernel-name K --print-source cuda,sass > k.csv
(the capture needs --import-source on and a -lineinfo build).
usage: python tools/src_hist.py k.csv [units] [top]   -- `units` divides the counts (e.g. strips)."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    fname, out, tot = '?', [], 0
    ie = isamp = None
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Name':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            ie, isamp = r.index('Instructions Executed'), r.index('# Samples')
            continue
        if ie is None or not r[0].isdigit() or len(r) <= ie or not r[ie].isdigit():
            continue
        n = int(r[ie])
        if n:
            out.append((n, int(r[isamp]) if r[isamp].isdigit() else 0, fname, int(r[0]), r[1]))
            tot += n
    print(f'total {tot}  per unit {tot / units:.1f}')
    for n, s, f, line, src in sorted(out, reverse=True)[:top]:
        print(f'{n / units:8.1f} {100.0 * n / tot:5.1f}% samp {s:4d}  {f}:{line}: {src.strip()[:100]}')


if __name__ == '__main__':
    main()
