"""TEST INFRASTRUCTURE ONLY - golden of model_mean_type='noise' from the UNMODIFIED reference (needs /root/reference):

    python -m oracle.make_golden_noise

The reference's own DecompScorePosNet3D.sample_diffusion with `model_mean_type: noise` (models/decompdiff.py:602-605: the network
output minus x_t is the predicted noise, x_0 comes from `_predict_x0_from_eps`) on a seeded synthetic batch, with and without the
drift guidance.  Inputs and weights are regenerated from their seeds on the GPU box; only the reference's outputs are committed.
"""
from __future__ import annotations

import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from decompdiff_b200 import synthetic as syn  # noqa: E402
from oracle import ref_shims  # noqa: E402
from oracle.make_golden import GOLDEN_DIR  # noqa: E402

NOISE_CASES = {
    'traj_noise_T12': dict(batch=dict(n_pockets=2, n_protein=150, arm_sizes=(5, 4), n_scaffold=8, seed=41), num_steps=12, noise_seed=411,
                           drift=None),
    'traj_noise_T6_guided': dict(batch=dict(n_pockets=3, n_protein=100, arm_sizes=(4, 4), n_scaffold=7, seed=42, n_full_extra=120),
                                 num_steps=6, noise_seed=412,
                                 drift=[{'type': 'armsca_prox', 'min_d': 1.2, 'max_d': 1.9}, {'type': 'clash', 'sigma': 2, 'gamma': 4}]),
}


def main():
    ref_shims.load_reference()
    from models.decompdiff import DecompScorePosNet3D
    cfg = ref_shims.reference_model_config()
    cfg.model_mean_type = 'noise'
    model = DecompScorePosNet3D(cfg, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES).eval()
    model.load_state_dict(syn.synthetic_state_dict(model, seed=0), strict=True)
    for name, spec in NOISE_CASES.items():
        batch = syn.make_batch(**spec['batch'])
        torch.manual_seed(spec['noise_seed'])      # the loop draws from the global CPU generator
        r = model.sample_diffusion(**batch, num_steps=spec['num_steps'], center_pos_mode='protein', energy_drift_opt=spec['drift'])
        gold = {'pos': r['pos'], 'v': r['v'], 'bond': r['bond'], 'pos_traj': torch.stack(r['pos_traj']),
                'v_traj': torch.stack(r['v_traj']).to(torch.int8), 'bond_traj': torch.stack(r['bond_traj']).to(torch.int8)}
        torch.save(gold, os.path.join(GOLDEN_DIR, f'{name}.pt'))
        print(name, tuple(gold['pos_traj'].shape), float(gold['pos'].abs().max()))


if __name__ == '__main__':
    main()
