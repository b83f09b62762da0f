"""Instruction mix and stall samples per SASS opcode from `ncu --page source --csv` (one or more kernels per file).
usage: python profiles/opmix2.py src.csv [units_per_launch]   (units: e.g. tiles, to print instructions per unit per warp)"""
import csv
import sys
from collections import defaultdict


def main(path, units=None):
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == 'Kernel Name':
            name = rows[i][1]
            hdr = rows[i + 1]
            si, st, ei = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
            j = i + 2
            ex, smp = defaultdict(int), defaultdict(int)
            while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
                r = rows[j]
                try:
                    n, e = int(r[st] or 0), int(r[ei] or 0)
                except Exception:
                    j += 1
                    continue
                toks = r[si].strip().split()
                op = toks[0] if toks and not toks[0].startswith('@') else (toks[1] if len(toks) > 1 else '?')
                op = '.'.join(op.split('.')[:2])
                ex[op] += e
                smp[op] += n
                j += 1
            tot_e, tot_s = sum(ex.values()), sum(smp.values()) or 1
            print(f'== {name}: {tot_e} warp instructions, {tot_s} samples' + (f', {tot_e / units:.0f} per unit' if units else ''))
            for op in sorted(ex, key=lambda o: -smp[o])[:28]:
                print(f'  {op:22s} exec {ex[op]:11d} ({ex[op] / tot_e:6.3f})' + (f' {ex[op] / units:8.1f}/unit' if units else '') +
                      f'  samples {smp[op] / tot_s:6.3f}')
            i = j
        else:
            i += 1


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
