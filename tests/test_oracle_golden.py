"""Pins the CPU oracle (oracle/restate.py) against outputs of the UNMODIFIED reference
(tests/golden/*.pt, written by oracle/make_golden.py in the build container)."""
import pytest
import torch

from conftest import load_golden, tol_ratio
from decompdiff_b200 import synthetic as syn
from oracle import fused_algebra, make_golden, make_golden_hybrid, make_golden_noise, make_golden_time, restate


@pytest.mark.parametrize('case', list(make_golden.FORWARD_CASES))
def test_oracle_forward_matches_reference(case, weights, oracle_cfg):
    kw = syn.make_batch(**make_golden.FORWARD_CASES[case])
    gold = load_golden(case)
    with torch.no_grad():
        out = restate.forward(weights, oracle_cfg, **syn.forward_kwargs(kw, None))
    for k in ('pred_ligand_pos', 'pred_ligand_v', 'pred_bond'):
        assert out[k].shape == gold[k].shape
        # bit-identical in the build container; other hosts / BLAS kernels may re-associate fp32 sums
        assert tol_ratio(out[k], gold[k]) <= 0.2, (case, k, tol_ratio(out[k], gold[k]))


@pytest.mark.parametrize('case', list(make_golden_hybrid.HYBRID_CASES))
def test_oracle_hybrid_mode_matches_reference(case, weights, oracle_cfg):
    """cutoff_mode='hybrid': the oracle's edge list is the reference's batch_hybrid_edge_connection edge for edge (same order), and
    its forward reproduces the reference model's outputs."""
    kw = syn.make_batch(**make_golden_hybrid.HYBRID_CASES[case])
    gold = load_golden(case)
    with torch.no_grad():
        out = restate.forward(weights, dict(oracle_cfg, cutoff_mode='hybrid'), **syn.forward_kwargs(kw, None), return_all=True)
    assert torch.equal(out['edge_index'].to(torch.int32), gold['edge_index'])
    for k in ('pred_ligand_pos', 'pred_ligand_v', 'pred_bond'):
        assert tol_ratio(out[k], gold[k]) <= 0.2, (case, k, tol_ratio(out[k], gold[k]))


def test_oracle_guided_trajectory_matches_reference(weights, oracle_cfg):
    spec = make_golden.TRAJ_CASES['traj_b3_T8_guided']
    kw = syn.make_batch(**spec['batch'])
    gold = load_golden('traj_b3_T8_guided')
    n, Eb = kw['init_ligand_pos'].size(0), kw['init_ligand_fc_bond_type'].numel()
    noise = syn.step_noise(n, Eb, spec['num_steps'], spec['noise_seed'])
    r = restate.sample_diffusion(weights, oracle_cfg, **kw, num_steps=spec['num_steps'], center_pos_mode='protein',
                                 energy_drift_opt=spec['drift'], noise=noise)
    assert torch.equal(torch.stack(r['v_traj']), gold['v_traj'].long())
    assert torch.equal(torch.stack(r['bond_traj']), gold['bond_traj'].long())
    assert tol_ratio(torch.stack(r['pos_traj']), gold['pos_traj']) <= 1.0
    assert tol_ratio(r['vt_traj'][-1], gold['vt_last']) <= 1.0


@pytest.mark.parametrize('case', list(make_golden_noise.NOISE_CASES))
def test_oracle_noise_mean_type_matches_reference(case, weights, oracle_cfg):
    """model_mean_type='noise' (models/decompdiff.py:602-605): the restatement's branch against the unmodified reference's run."""
    spec = make_golden_noise.NOISE_CASES[case]
    kw = syn.make_batch(**spec['batch'])
    gold = load_golden(case)
    n, Eb = kw['init_ligand_pos'].size(0), kw['init_ligand_fc_bond_type'].numel()
    noise = syn.step_noise(n, Eb, spec['num_steps'], spec['noise_seed'])
    cfg = dict(oracle_cfg, model_mean_type='noise')
    r = restate.sample_diffusion(weights, cfg, **kw, num_steps=spec['num_steps'], center_pos_mode='protein',
                                 energy_drift_opt=spec['drift'], noise=noise)
    assert torch.equal(torch.stack(r['v_traj']), gold['v_traj'].long())
    assert torch.equal(torch.stack(r['bond_traj']), gold['bond_traj'].long())
    assert tol_ratio(torch.stack(r['pos_traj']), gold['pos_traj']) <= 1.0
    c0 = restate.sample_diffusion(weights, oracle_cfg, **kw, num_steps=2, center_pos_mode='protein', energy_drift_opt=spec['drift'], noise=noise)
    assert float((c0['pos_traj'][1] - gold['pos_traj'][1]).abs().max()) > 1e-3      # the branch changes the trajectory


def _time_weights():
    """Name-keyed synthetic weights of the model WITH the 'simple' time embedding (ligand_atom_emb has one more input column)."""
    import decompdiff_b200 as ddb
    m = ddb.DecompScorePosNet3D(dict(syn.DEFAULT_MODEL_CONFIG, time_emb_dim=1, time_emb_mode='simple'), syn.PROTEIN_FEATURE_DIM,
                                syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
    m.load_state_dict(syn.synthetic_state_dict(m, seed=0))
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def test_oracle_simple_time_embedding_matches_reference(oracle_cfg):
    """time_emb_dim > 0, time_emb_mode='simple' (models/decompdiff.py:224-229): forward with one time step per graph and a short
    sampling run against the unmodified reference."""
    _, w = _time_weights()
    assert w['ligand_atom_emb.weight'].shape[1] == syn.LIGAND_FEATURE_DIM + 1
    cfg = dict(oracle_cfg, time_emb_dim=1, time_emb_mode='simple')
    kw = syn.make_batch(**make_golden_time.TIME_FWD['batch'])
    gold = load_golden('fwd_time_simple')
    with torch.no_grad():
        out = restate.forward(w, cfg, **syn.forward_kwargs(kw, torch.tensor(make_golden_time.TIME_FWD['time_step'])))
    for k in ('pred_ligand_pos', 'pred_ligand_v', 'pred_bond'):
        assert tol_ratio(out[k], gold[k]) <= 0.2, (k, tol_ratio(out[k], gold[k]))
    assert float((gold['pred_ligand_v'] - gold['pred_ligand_v_t500']).abs().max()) > 1e-2      # the time step matters
    spec = make_golden_time.TIME_TRAJ
    kw = syn.make_batch(**spec['batch'])
    gold = load_golden('traj_time_simple')
    n, Eb = kw['init_ligand_pos'].size(0), kw['init_ligand_fc_bond_type'].numel()
    noise = syn.step_noise(n, Eb, spec['num_steps'], spec['noise_seed'])
    r = restate.sample_diffusion(w, cfg, **kw, num_steps=spec['num_steps'], center_pos_mode='protein', noise=noise)
    assert torch.equal(torch.stack(r['v_traj']), gold['v_traj'].long()) and torch.equal(torch.stack(r['bond_traj']), gold['bond_traj'].long())
    assert tol_ratio(torch.stack(r['pos_traj']), gold['pos_traj']) <= 1.0


def test_oracle_cfg1_first_steps_match_reference(weights, oracle_cfg):
    spec = make_golden.TRAJ_CASES['traj_cfg1_T50']
    kw = syn.make_batch(**spec['batch'])
    gold = load_golden('traj_cfg1_T50')
    steps = 4     # the loop starts at t = T-1 whatever num_steps is, so a prefix of the golden trajectory is comparable
    noise = syn.step_noise(30, 870, steps, spec['noise_seed'])
    r = restate.sample_diffusion(weights, oracle_cfg, **kw, num_steps=steps, center_pos_mode='protein', noise=noise)
    assert torch.equal(torch.stack(r['v_traj']), gold['v_traj'][:steps].long())
    assert torch.equal(torch.stack(r['bond_traj']), gold['bond_traj'][:steps].long())
    assert tol_ratio(torch.stack(r['pos_traj']), gold['pos_traj'][:steps]) <= 1.0
    assert tol_ratio(r['v0_traj'][0], gold['v0_first']) <= 1.0 and tol_ratio(r['bt_traj'][0], gold['bt_first']) <= 1.0


def test_oracle_guidance_gradients_match_reference():
    spec = make_golden.TRAJ_CASES['traj_b3_T8_guided']
    kw = syn.make_batch(**spec['batch'])
    gold = load_golden('guidance_grads')
    x = kw['init_ligand_pos']
    zero = torch.zeros_like(x)
    g1 = restate.guidance_grad(x, zero, [spec['drift'][0]], kw['batch_ligand'], kw['ligand_decomp_index'])
    g2 = restate.guidance_grad(x, zero, [spec['drift'][1]], kw['batch_ligand'], kw['ligand_decomp_index'],
                               kw['full_protein_pos'], kw['full_batch_protein'])
    assert tol_ratio(g1, gold['armsca_grad']) <= 1.0 and tol_ratio(g2, gold['clash_grad']) <= 1.0


def test_fused_algebra_within_tolerance(weights, oracle_cfg):
    """The re-associations the CUDA kernels rely on (first-Linear split, key / value contractions) keep the
    result inside rtol 1e-4 / atol 1e-5 of the reference formulation."""
    kw = syn.make_batch(n_pockets=3, n_protein=120, arm_sizes=(5, 4), n_scaffold=8, seed=31, ragged=True)
    fk = syn.forward_kwargs(kw, None)
    with torch.no_grad():
        ref = restate.forward(weights, oracle_cfg, **fk)
        fused = fused_algebra.forward_fused(weights, oracle_cfg, **fk)
    for k in ref:
        assert tol_ratio(fused[k], ref[k]) <= 0.25, (k, tol_ratio(fused[k], ref[k]))


def test_triplet_enumeration_counts():
    """n(n-1)(n-2) triplets for a fully connected ligand; groups ordered by the j->i edge (uni_transformer_edge.py:103-123)."""
    bi = syn.fc_bond_index(6)
    idx_i, idx_j, idx_k, idx_kj, idx_ji = restate.bond_triplets(bi[0], bi[1], 6)
    assert idx_i.numel() == 6 * 5 * 4
    assert bool((idx_i != idx_k).all()) and bool((idx_j != idx_k).all()) and bool((idx_i != idx_j).all())
    assert bool((bi[0][idx_kj] == idx_k).all()) and bool((bi[1][idx_kj] == idx_j).all())
    assert bool((bi[0][idx_ji] == idx_j).all()) and bool((bi[1][idx_ji] == idx_i).all())
    assert bool((idx_ji[1:] >= idx_ji[:-1]).all())
