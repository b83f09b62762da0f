"""Opcode mix / hottest instructions of one kernel from `ncu --page source --csv` output."""
import collections
import csv
import sys


def main(path, top=22):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    si, ei, st = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    agg, samp = collections.Counter(), collections.Counter()
    tot = tots = 0
    for r in rows[2:]:
        try:
            n, s = int(r[ei]), int(r[st])
        except Exception:
            continue
        toks = r[si].strip().split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = op.split('.')[0]
        agg[op] += n; samp[op] += s; tot += n; tots += s
    print(f'{"op":10s} {"executed":>14s} {"share":>7s} {"stall-sample share":>18s}')
    for k, v in agg.most_common(top):
        print(f'{k:10s} {v:14d} {v / tot:7.3f} {samp[k] / max(tots, 1):18.3f}')
    print('total warp instructions', tot)


if __name__ == '__main__':
    main(sys.argv[1])
