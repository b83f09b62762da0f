"""Where ptxas spills: STL / LDL instructions of a kernel by source line.  usage: python profiles/spillmap.py file.o|file.cubin <kernel substr>"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main(obj, want):
    with tempfile.TemporaryDirectory() as tmp:
        cubin = obj
        if not obj.endswith('.cubin'):
            subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
            cubin = os.path.join(tmp, sorted(os.listdir(tmp))[0])
        sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], check=True, capture_output=True, text=True).stdout
    cur, on, cnt = None, False, collections.Counter()
    for ln in sass.splitlines():
        if ln.startswith('.text.'):
            on = want in ln
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if on:
            m = re.search(r'\b(STL|LDL)(\.\d+)?\b', ln)
            if m:
                cnt[(cur, m.group(1))] += 1
    for k, v in sorted(cnt.items()):
        print(k, v)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
