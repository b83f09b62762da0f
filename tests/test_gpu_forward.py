"""GPU parity of DecompScorePosNet3D.forward: CUDA path (through the C ABI) vs the committed reference
outputs (tests/golden, produced by the unmodified reference) and vs the CPU oracle on the same inputs."""
import pytest
import torch

from conftest import load_golden, tol_ratio
from decompdiff_b200 import synthetic as syn
from oracle import make_golden, restate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(make_golden.FORWARD_CASES))
def test_forward_matches_reference_golden(case, model_cpu):
    kw = syn.make_batch(**make_golden.FORWARD_CASES[case])
    gold = load_golden(case)
    out = model_cpu(**syn.forward_kwargs(kw, None))
    for k in ('pred_ligand_pos', 'pred_ligand_v', 'pred_bond'):
        assert out[k].shape == gold[k].shape
        r = tol_ratio(out[k], gold[k])
        print(case, k, 'max err / tol =', r)
        assert r <= 1.0, f'{case}:{k} outside rtol 1e-4 / atol 1e-5 ({r:.3f} x tol)'


def test_forward_matches_oracle_new_inputs(model_cpu, weights, oracle_cfg):
    kw = syn.make_batch(n_pockets=4, n_protein=150, arm_sizes=(5, 6), n_scaffold=9, seed=77, ragged=True)
    fk = syn.forward_kwargs(kw, None)
    out = model_cpu(**fk)
    with torch.no_grad():
        ref = restate.forward(weights, oracle_cfg, **fk)
    for k in ('pred_ligand_pos', 'pred_ligand_v', 'pred_bond'):
        assert tol_ratio(out[k], ref[k]) <= 1.0
