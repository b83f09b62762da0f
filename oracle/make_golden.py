"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference's own `DecompScorePosNet3D` is imported through `oracle/ref_shims.py`, loaded with the
name-keyed synthetic weights (`decompdiff_b200.synthetic.synthetic_state_dict`, seed 0) and evaluated on
seeded synthetic pockets.  Inputs and weights are NOT stored - they are regenerated from their seeds on the
GPU box; only the reference's outputs are committed (a few hundred KB).
"""
from __future__ import annotations

import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from decompdiff_b200 import synthetic as syn  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, 'tests', 'golden')

# name -> make_batch kwargs (kept tiny so the fixtures stay small and the oracle runs in seconds)
FORWARD_CASES = {
    'fwd_cfg1': dict(n_pockets=1, n_protein=300, arm_sizes=(8, 8), n_scaffold=14, seed=11),          # BASELINE cfg 1 shape
    'fwd_b2_400': dict(n_pockets=2, n_protein=370, arm_sizes=(8, 8), n_scaffold=14, seed=12),        # cfg 2 per-pocket shape
    'fwd_ragged': dict(n_pockets=5, n_protein=90, arm_sizes=(4, 5), n_scaffold=7, seed=13, ragged=True),
    'fwd_tiny_graphs': dict(n_pockets=3, n_protein=[9, 20, 33], arm_sizes=(2,), n_scaffold=3, seed=14),  # < k+1 nodes
    'fwd_dense': dict(n_pockets=2, n_protein=200, arm_sizes=(6, 6, 5), n_scaffold=10, seed=15, dense=True),
}
TRAJ_CASES = {
    'traj_cfg1_T50': dict(batch=dict(n_pockets=1, n_protein=300, arm_sizes=(8, 8), n_scaffold=14, seed=21),
                          num_steps=50, noise_seed=2021, drift=None),
    'traj_b3_T8_guided': dict(batch=dict(n_pockets=3, n_protein=120, arm_sizes=(5, 4), n_scaffold=8, seed=22,
                                         n_full_extra=150),
                              num_steps=8, noise_seed=7,
                              drift=[{'type': 'armsca_prox', 'min_d': 1.2, 'max_d': 1.9},
                                     {'type': 'clash', 'sigma': 2, 'gamma': 4}]),
}


def reference_model():
    ref_shims.load_reference()
    from models.decompdiff import DecompScorePosNet3D
    cfg = ref_shims.reference_model_config()
    model = DecompScorePosNet3D(cfg, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES).eval()
    model.load_state_dict(syn.synthetic_state_dict(model, seed=0), strict=True)
    return model


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    model = reference_model()
    keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    torch.save(keys, os.path.join(GOLDEN_DIR, 'state_dict_keys.pt'))
    for name, kw in FORWARD_CASES.items():
        batch = syn.make_batch(**kw)
        with torch.no_grad():
            out = model(**syn.forward_kwargs(batch, torch.full((kw['n_pockets'],), 500)))
        torch.save({k: v.clone() for k, v in out.items()}, os.path.join(GOLDEN_DIR, f'{name}.pt'))
        print(name, {k: tuple(v.shape) for k, v in out.items()})
    for name, spec in TRAJ_CASES.items():
        batch = syn.make_batch(**spec['batch'])
        torch.manual_seed(spec['noise_seed'])      # the loop draws from the global CPU generator
        r = model.sample_diffusion(**batch, num_steps=spec['num_steps'], center_pos_mode='protein',
                                   energy_drift_opt=spec['drift'])
        gold = {'pos': r['pos'], 'v': r['v'], 'bond': r['bond'],
                'pos_traj': torch.stack(r['pos_traj']), 'v_traj': torch.stack(r['v_traj']).to(torch.int8),
                'bond_traj': torch.stack(r['bond_traj']).to(torch.int8),
                'vt_last': r['vt_traj'][-1], 'v0_first': r['v0_traj'][0], 'bt_first': r['bt_traj'][0]}
        torch.save(gold, os.path.join(GOLDEN_DIR, f'{name}.pt'))
        print(name, tuple(gold['pos_traj'].shape))
    # guidance gradients of the reference's own energy functions (utils/guidance_funcs.py)
    import utils.guidance_funcs as guidance
    spec = TRAJ_CASES['traj_b3_T8_guided']
    batch = syn.make_batch(**spec['batch'])
    x = batch['init_ligand_pos'].clone().requires_grad_(True)
    e1, _ = guidance.compute_batch_armsca_prox_loss(x, batch['batch_ligand'], batch['ligand_decomp_index'], min_d=1.2, max_d=1.9)
    g1 = torch.autograd.grad(e1, x)[0]
    e2 = guidance.compute_batch_clash_loss(batch['full_protein_pos'], x, batch['full_batch_protein'], batch['batch_ligand'],
                                           sigma=2, surface_ct=4)
    g2 = torch.autograd.grad(e2, x)[0]
    torch.save({'armsca_grad': g1, 'clash_grad': g2, 'armsca_e': e1.detach(), 'clash_e': e2.detach()},
               os.path.join(GOLDEN_DIR, 'guidance_grads.pt'))
    print('guidance', float(g1.abs().max()), float(g2.abs().max()))


if __name__ == '__main__':
    main()
