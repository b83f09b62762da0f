"""TEST INFRASTRUCTURE ONLY - goldens of cutoff_mode='hybrid' from the UNMODIFIED reference (needs /root/reference):

    python -m oracle.make_golden_hybrid

`batch_hybrid_edge_connection` (models/common.py:250-277: ligand-ligand fully connected, k nearest protein atoms per ligand atom,
kNN over all atoms for the protein destinations) is exercised twice: directly on seeded coordinates (edge list fixture) and through
the reference's own DecompScorePosNet3D.forward with `cutoff_mode: hybrid`.  Inputs and weights are regenerated from their seeds
on the GPU box; only the reference's outputs are committed.
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

HYBRID_CASES = {
    'fwd_hybrid': dict(n_pockets=2, n_protein=150, arm_sizes=(6, 5), n_scaffold=9, seed=31),               # 20-atom ligands: 19 + 32 edges per ligand atom
    'fwd_hybrid_ragged': dict(n_pockets=4, n_protein=80, arm_sizes=(3, 4), n_scaffold=5, seed=32, ragged=True),
    'fwd_hybrid_large': dict(n_pockets=1, n_protein=200, arm_sizes=(12, 12), n_scaffold=16, seed=33),       # 40-atom ligand: 71 edges
}


def main():
    ref_shims.load_reference()
    from models.common import batch_hybrid_edge_connection
    from oracle import restate
    from models.decompdiff import DecompScorePosNet3D
    cfg = ref_shims.reference_model_config()
    cfg.cutoff_mode = 'hybrid'
    model = DecompScorePosNet3D(cfg, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES).eval()
    model.load_state_dict(syn.synthetic_state_dict(model, seed=0), strict=True)
    for name, kw in HYBRID_CASES.items():
        batch = syn.make_batch(**kw)
        n_pockets = kw['n_pockets']
        with torch.no_grad():
            out = model(**syn.forward_kwargs(batch, torch.full((n_pockets,), 500)))
        gold = {k: v.clone() for k, v in out.items()}
        # the edge list of the same batch, in the merged node order of compose_context (common.py:172-191)
        hp = torch.zeros(batch['protein_pos'].size(0), 1)
        hl = torch.ones(batch['init_ligand_pos'].size(0), 1)
        _, x, b_all, mask_l, _, _ = restate.compose_context(hp, hl, batch['protein_pos'], batch['init_ligand_pos'], batch['batch_protein'],
                                                            batch['batch_ligand'])
        gold['edge_index'] = batch_hybrid_edge_connection(x, k=cfg.knn, mask_ligand=mask_l, batch=b_all, add_p_index=True).to(torch.int32)
        torch.save(gold, os.path.join(GOLDEN_DIR, f'{name}.pt'))
        print(name, {k: tuple(v.shape) for k, v in gold.items()})


if __name__ == '__main__':
    main()
