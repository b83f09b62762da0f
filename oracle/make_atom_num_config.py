"""TEST / DATA INFRASTRUCTURE - export the reference's empirical atom-count table to decompdiff_b200/data/atom_num_config.json.

`utils/evaluation/atom_num_config.py` (CONFIG) is a data table (bin bounds of the pocket "space size" and, per bin, the observed
atom counts with their frequencies); `sample_atom_num` (utils/evaluation/atom_num.py:20-35) always bins with CONFIG['bounds']
and falls back to CONFIG['bins'] when no dictionary is passed.  The product needs the same numbers to reproduce the
reference's draws, so the table is exported once as JSON:   python -m oracle.make_atom_num_config   (needs /root/reference)
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sys.path.insert(0, '/root/reference')
    from utils.evaluation.atom_num_config import CONFIG
    out = {'source': 'bytedance/DecompDiff utils/evaluation/atom_num_config.py (CC-BY-NC-4.0), data table only',
           'bounds': [float(b) for b in CONFIG['bounds']],
           'bins': [[[int(c) for c in counts], [float(p) for p in probs]] for counts, probs in CONFIG['bins']]}
    path = os.path.join(REPO, 'decompdiff_b200', 'data', 'atom_num_config.json')
    with open(path, 'w') as f:
        json.dump(out, f)
    print(path, len(out['bins']), 'bins')


if __name__ == '__main__':
    main()
