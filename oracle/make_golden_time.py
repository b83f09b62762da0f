"""TEST INFRASTRUCTURE ONLY - goldens of the 'simple' time embedding from the UNMODIFIED reference (needs /root/reference):

    python -m oracle.make_golden_time

The reference's own DecompScorePosNet3D with `time_emb_dim: 1, time_emb_mode: simple` (models/decompdiff.py:168-173, 224-229: the
ligand feature vector gets the column time_step / num_timesteps): a forward with a different time step per graph and a short
sample_diffusion run.  Inputs and weights are regenerated from their seeds on the GPU box; only the reference's outputs are stored.
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

TIME_FWD = dict(batch=dict(n_pockets=4, n_protein=120, arm_sizes=(5, 4), n_scaffold=7, seed=51, ragged=True), time_step=[999, 640, 17, 0])
TIME_TRAJ = dict(batch=dict(n_pockets=2, n_protein=150, arm_sizes=(5, 4), n_scaffold=8, seed=52), num_steps=8, noise_seed=521)


def main():
    ref_shims.load_reference()
    from models.decompdiff import DecompScorePosNet3D
    cfg = ref_shims.reference_model_config()
    cfg.time_emb_dim, cfg.time_emb_mode = 1, 'simple'
    model = DecompScorePosNet3D(cfg, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES).eval()
    model.load_state_dict(syn.synthetic_state_dict(model, seed=0), strict=True)
    batch = syn.make_batch(**TIME_FWD['batch'])
    with torch.no_grad():
        out = model(**syn.forward_kwargs(batch, torch.tensor(TIME_FWD['time_step'])))
        other = model(**syn.forward_kwargs(batch, torch.full((4,), 500)))
    gold = {k: v.clone() for k, v in out.items()}
    gold['pred_ligand_v_t500'] = other['pred_ligand_v'].clone()
    torch.save(gold, os.path.join(GOLDEN_DIR, 'fwd_time_simple.pt'))
    print('fwd_time_simple', float((out['pred_ligand_v'] - other['pred_ligand_v']).abs().max()))
    batch = syn.make_batch(**TIME_TRAJ['batch'])
    torch.manual_seed(TIME_TRAJ['noise_seed'])
    r = model.sample_diffusion(**batch, num_steps=TIME_TRAJ['num_steps'], center_pos_mode='protein')
    gold = {'pos': r['pos'], 'v': r['v'], 'bond': r['bond'], 'pos_traj': torch.stack(r['pos_traj']),
            'v_traj': torch.stack(r['v_traj']).to(torch.int8), 'bond_traj': torch.stack(r['bond_traj']).to(torch.int8)}
    torch.save(gold, os.path.join(GOLDEN_DIR, 'traj_time_simple.pt'))
    print('traj_time_simple', tuple(gold['pos_traj'].shape))


if __name__ == '__main__':
    main()
