"""Stall samples aggregated per CUDA source line from `ncu --page source --csv --print-source cuda,sass`."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file = '?'
    agg = collections.Counter()
    text = {}
    hdr = None
    tot = 0
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            si = hdr.index('# Samples')
            continue
        if hdr is None or r[0] in ('Function Name', 'File Name'):
            continue
        try:
            line = int(r[0]) if r[0] else None
        except ValueError:
            continue
        if line is not None:
            cur_line = (cur_file, line)
            text[cur_line] = r[1].strip()[:90]
        try:
            n = int(r[si])
        except Exception:
            continue
        if r[2]:      # a SASS row (has an address)
            agg[cur_line] += n
            tot += n
    print('total samples', tot)
    for (f, l), n in agg.most_common(top):
        print(f'{n / tot:6.3f} {f}:{l:4d}  {text.get((f, l), "")}')


if __name__ == '__main__':
    main(sys.argv[1])
