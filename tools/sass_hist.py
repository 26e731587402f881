#!/usr/bin/env python
"""Histogram of executed SASS instructions per opcode from an `ncu --page source --csv` export.

usage: ncu -i X.ncu-rep --page source --csv --kernel-name K > k.csv; python tools/sass_hist.py k.csv [units]
`units` divides the counts (e.g. the number of strips) to give instructions per unit of work."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hdr = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
    h = rows[hdr]
    ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for row in rows[hdr + 1:]:
        if row and row[0] == 'Address':        # a second section (CUDA source lines) follows the SASS one
            break
        if len(row) <= ie or not row[ie].isdigit():
            continue
        t = row[ia].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]
        n = int(row[ie])
        ops[op] += n
        samp[op] += int(row[isamp] or 0)
        tot += n
    print(f'total warp instructions {tot}  per unit {tot / units:.1f}')
    for op, n in ops.most_common(40):
        print(f'{op:12s} {n:10d} {n / units:8.1f}  {100.0 * n / tot:5.1f}%  stall samples {samp[op]}')


if __name__ == '__main__':
    main()
