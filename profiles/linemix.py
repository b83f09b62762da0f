"""Executed instructions and stall samples per SOURCE LINE: joins `ncu --page source --csv` (per SASS instruction, in program order)
with `nvdisasm -g -c` of the same cubin (line info).  usage: python profiles/linemix.py src.csv file.sass <kernel substr in csv> <kernel substr in sass> [units]"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(path, want):
    out, cur, on = [], None, False
    for ln in open(path):
        if ln.startswith('.text.'):
            on = want in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main(src_csv, sass, want_csv, want, units=None):
    rows = list(csv.reader(open(src_csv)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == 'Kernel Name' and want_csv in rows[i][1]:
            break
        i += 1
    hdr = rows[i + 1]
    st, ei = hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [(k, c) for k, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    data, reasons = [], []
    j = i + 2
    while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
        try:
            data.append((int(rows[j][st] or 0), int(rows[j][ei] or 0)))
            reasons.append({c[6:]: int(rows[j][k] or 0) for k, c in stall_cols})
        except Exception:
            pass
        j += 1
    sl = sass_lines(sass, want)
    print(len(data), 'profiled instructions,', len(sl), 'disassembled')
    n = min(len(data), len(sl))
    ex, sm, why, total_why = defaultdict(int), defaultdict(int), defaultdict(lambda: defaultdict(int)), defaultdict(int)
    for (s_, e_), (line, _), rs in zip(data[:n], sl[:n], reasons[:n]):
        ex[line] += e_
        sm[line] += s_
        for k, v in rs.items():
            why[line][k] += v
            total_why[k] += v
    te, ts = sum(ex.values()), sum(sm.values()) or 1
    for line in sorted(ex, key=lambda l: -sm[l])[:45]:
        top = sorted(why[line].items(), key=lambda kv: -kv[1])[:2]
        print(f'{str(line):38s} exec {ex[line] / te:6.3f}' + (f' {ex[line] / units:7.1f}/unit' if units else '') + f'  samples {sm[line] / ts:6.3f}  '
              + ' '.join(f'{k}={v / max(sm[line], 1):.2f}' for k, v in top if v))
    tw = sum(total_why.values()) or 1
    print('stall reasons overall: ' + ' '.join(f'{k}={v / tw:.3f}' for k, v in sorted(total_why.items(), key=lambda kv: -kv[1])[:10]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]) if len(sys.argv) > 5 else None)
