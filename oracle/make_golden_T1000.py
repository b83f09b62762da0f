"""TEST INFRASTRUCTURE ONLY - the full T=1000 cfg-1 trajectory of the UNMODIFIED reference (tests/golden/traj_cfg1_T1000.pt).

Run in the build container (needs /root/reference; ~6 min on 8 host threads):   python -m oracle.make_golden_T1000
Same recipe as `oracle/make_golden.py` (reference model through `oracle/ref_shims.py`, name-keyed synthetic weights, seeded
synthetic pocket, the reference's own noise stream from `torch.manual_seed`).  Discrete samples are stored as int8.
"""
from __future__ import annotations

import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from decompdiff_b200 import synthetic as syn  # noqa: E402
from oracle import make_golden  # noqa: E402

CASE = dict(batch=dict(n_pockets=1, n_protein=300, arm_sizes=(8, 8), n_scaffold=14, seed=21), num_steps=1000, noise_seed=2021)


def main():
    model = make_golden.reference_model()
    batch = syn.make_batch(**CASE['batch'])
    torch.manual_seed(CASE['noise_seed'])
    t0 = time.time()
    r = model.sample_diffusion(**batch, num_steps=CASE['num_steps'], center_pos_mode='protein', energy_drift_opt=None)
    print('reference T=1000 run: %.1f s' % (time.time() - t0))
    gold = {'pos': r['pos'], 'v': r['v'], 'bond': r['bond'],
            'pos_traj': torch.stack(r['pos_traj']), 'v_traj': torch.stack(r['v_traj']).to(torch.int8),
            'bond_traj': torch.stack(r['bond_traj']).to(torch.int8)}
    torch.save(gold, os.path.join(make_golden.GOLDEN_DIR, 'traj_cfg1_T1000.pt'))
    print({k: tuple(v.shape) for k, v in gold.items()})


if __name__ == '__main__':
    main()
