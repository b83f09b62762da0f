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
    data = []
    j = i + 2
    while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
        try:
            data.append((int(rows[j][st] or 0), int(rows[j][ei] or 0)))
        except Exception:
            pass
        j += 1
    sl = sass_lines(sass, want)
    print(len(data), 'profiled instructions,', len(sl), 'disassembled')
    n = min(len(data), len(sl))
    ex, sm = defaultdict(int), defaultdict(int)
    for (s_, e_), (line, _) in zip(data[:n], sl[:n]):
        ex[line] += e_
        sm[line] += s_
    te, ts = sum(ex.values()), sum(sm.values()) or 1
    for line in sorted(ex, key=lambda l: -sm[l])[:45]:
        print(f'{str(line):38s} exec {ex[line] / te:6.3f}' + (f' {ex[line] / units:7.1f}/unit' if units else '') + f'  samples {sm[line] / ts:6.3f}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]) if len(sys.argv) > 5 else None)
