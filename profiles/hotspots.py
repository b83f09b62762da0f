"""Top stall-sample instructions of a kernel from `ncu --page source --csv` output (needs -lineinfo for file:line)."""
import csv
import sys


def main(path, top=28):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    si, st, ei = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    reasons = [i for i, h in enumerate(hdr) if h.startswith('stall_') and '(Not Issued)' not in h]
    data = []
    tot = 0
    for r in rows[2:]:
        try:
            n = int(r[st])
        except Exception:
            continue
        tot += n
        top_reason = max(reasons, key=lambda i: int(r[i] or 0))
        data.append((n, r[si].strip()[:70], hdr[top_reason], int(r[ei] or 0)))
    data.sort(reverse=True)
    print('total samples', tot)
    for n, src, why, ex in data[:top]:
        print(f'{n / tot:6.3f} {why:22s} exec={ex:10d} {src}')
    agg = {}
    for r in rows[2:]:
        for i in reasons:
            try:
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
            except Exception:
                pass
    s = sum(agg.values()) or 1
    print('stall mix:', ', '.join(f'{k[6:]}={v / s:.2f}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == '__main__':
    main(sys.argv[1])
