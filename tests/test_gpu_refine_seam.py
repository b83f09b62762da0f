"""The refine-net seam (contract 2): `get_refine_net('uni_o2_bond', config).forward(h, x, group_idx, bond_index,
h_bond, mask_ligand, mask_ligand_atom, batch)` on the CUDA kernels against the oracle's restatement of
UniTransformerO2TwoUpdateGeneralBond.forward (uni_transformer_edge.py:394-443) - final outputs for EVERY node, per-layer
intermediates, and `return_all` of the model API (decompdiff.py:343-350)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import tol_ratio
from decompdiff_b200 import synthetic as syn
from decompdiff_b200.decompdiff import as_config, get_refine_net
from decompdiff_b200.engine import RefineBatch
from oracle import restate

pytestmark = pytest.mark.gpu


def _merged_inputs(sd, kw):
    """h / x / bond_index / h_bond / masks / batch exactly as DecompScorePosNet3D.forward hands them to the refine net."""
    lig_feat = torch.cat([F.one_hot(kw['init_ligand_v'], 8).float(), kw['ligand_v_aux']], -1)
    h_p = F.linear(kw['protein_v'], sd['protein_atom_emb.weight'], sd['protein_atom_emb.bias'])
    h_l = F.linear(lig_feat, sd['ligand_atom_emb.weight'], sd['ligand_atom_emb.bias'])
    h_p = torch.cat([h_p, torch.zeros(h_p.size(0), 1)], -1)
    h_l = torch.cat([h_l, torch.ones(h_l.size(0), 1)], -1)
    h, x, batch, mask_l, mask_la, l_idx = restate.compose_context(h_p, h_l, kw['protein_pos'], kw['init_ligand_pos'], kw['batch_protein'],
                                                                 kw['batch_ligand'])
    bond_index = l_idx[kw['ligand_fc_bond_index']]
    h_bond = F.linear(F.one_hot(kw['init_ligand_fc_bond_type'], 5).float(), sd['ligand_bond_emb.weight'], sd['ligand_bond_emb.bias'])
    return h, x, bond_index, h_bond, mask_l, mask_la, batch


@pytest.mark.parametrize('batch_kw', [dict(n_pockets=2, n_protein=120, arm_sizes=(4, 5), n_scaffold=7, seed=81),
                                      dict(n_pockets=3, n_protein=[30, 200, 90], arm_sizes=(3,), n_scaffold=4, seed=82)])
def test_refine_net_forward_and_per_layer_intermediates(batch_kw, weights, oracle_cfg):
    kw = syn.make_batch(**batch_kw)
    h, x, bond_index, h_bond, mask_l, mask_la, batch = _merged_inputs(weights, kw)
    net = get_refine_net('uni_o2_bond', as_config(syn.DEFAULT_MODEL_CONFIG))
    net.load_state_dict({k[len('refine_net.'):]: v for k, v in weights.items() if k.startswith('refine_net.')})
    out = net(h.cuda(), x.cuda(), None, bond_index.cuda(), h_bond.cuda(), mask_l.cuda(), mask_la.cuda(), batch.cuda(), return_all=True)
    with torch.no_grad():
        ref = restate.refine_net(weights, oracle_cfg, h, x, bond_index, h_bond, mask_l, mask_la, batch, return_all=True)
    for k in ('x', 'h', 'h_bond'):
        assert out[k].is_cuda and out[k].shape == ref[k].shape
        r = tol_ratio(out[k], ref[k])
        print('refine', k, 'max err / tol', round(r, 4))
        assert r <= 1.0, k
    assert len(out['all_x']) == 2 and torch.equal(out['all_h'][0].cpu(), h) and torch.equal(out['all_x'][1], out['x'])
    # per-layer h / x / h_bond of every node (no pruning on this path) against the oracle's per-layer list
    rb = RefineBatch(net._refine_engine(torch.device('cuda', torch.cuda.current_device())), batch, mask_l, mask_la, bond_index)
    L = oracle_cfg['num_layers']
    _, _, _, (th, tx, thb) = rb.forward(h, x, h_bond, tap_layers=L)
    worst = 0.0
    for l in range(L):
        worst = max(worst, tol_ratio(th[l], ref['all_h'][l + 1]), tol_ratio(tx[l][:, :3], ref['all_x'][l + 1]),
                    tol_ratio(thb[l], ref['all_h_bond'][l + 1]))
    print('per-layer worst err / tol', round(worst, 4))
    assert worst <= 1.0


def test_refine_net_rejects_unsupported_layouts(weights):
    kw = syn.make_batch(n_pockets=1, n_protein=20, arm_sizes=(2,), n_scaffold=2, seed=83)
    h, x, bond_index, h_bond, mask_l, mask_la, batch = _merged_inputs(weights, kw)
    net = get_refine_net('uni_o2_bond', as_config(syn.DEFAULT_MODEL_CONFIG))
    net.load_state_dict({k[len('refine_net.'):]: v for k, v in weights.items() if k.startswith('refine_net.')})
    bad = mask_l.flip(0)      # ligand nodes ahead of protein nodes: not the compose_context order
    with pytest.raises(ValueError):
        net(h, x, None, bond_index, h_bond, bad, bad, batch)
    with pytest.raises(ValueError):
        net(h, x, None, None, None, mask_l, mask_la, batch)
    with pytest.raises(ValueError):
        get_refine_net('uni_o2', as_config(syn.DEFAULT_MODEL_CONFIG))


def test_forward_return_all(model_cpu, weights, oracle_cfg):
    kw = syn.make_batch(n_pockets=2, n_protein=80, arm_sizes=(4, 4), n_scaffold=5, seed=84)
    fk = syn.forward_kwargs(kw, None)
    out = model_cpu(**fk, return_all=True)
    assert len(out['layer_pred_ligand_pos']) == 2 and len(out['layer_pred_ligand_v']) == 2
    assert torch.equal(out['layer_pred_ligand_pos'][0], kw['init_ligand_pos']) and torch.equal(out['layer_pred_ligand_pos'][1], out['pred_ligand_pos'])
    lig_feat = torch.cat([F.one_hot(kw['init_ligand_v'], 8).float(), kw['ligand_v_aux']], -1)
    h_l = torch.cat([F.linear(lig_feat, weights['ligand_atom_emb.weight'], weights['ligand_atom_emb.bias']), torch.ones(lig_feat.size(0), 1)], -1)
    want = F.linear(restate.shifted_softplus(F.linear(h_l, weights['v_inference.0.weight'], weights['v_inference.0.bias'])),
                    weights['v_inference.2.weight'], weights['v_inference.2.bias'])
    assert tol_ratio(out['layer_pred_ligand_v'][0], want) <= 1.0
    assert torch.equal(out['layer_pred_ligand_v'][1], out['pred_ligand_v'])
