"""GPU parity at the shapes bench.py measures (BASELINE.json cfg 2 / 3 / 5) and at the ligand sizes SURVEY.md section 5 flags.

The CPU oracle needs ~1 s per 400-atom pocket and forward pass, so a 64-pocket batch is checked on a few of its pockets:
complexes never interact (per-graph kNN, per-graph scatters), hence the oracle on pockets {i} of the batch
(`synthetic.select_pockets`) must reproduce the rows of the full-batch CUDA result that belong to those pockets.  At these
sizes every persistent kernel walks ~40x more tiles per CTA than in the small goldens, and the level-sorted destination lists
and the first-layer cache lists are long."""
import pytest
import torch

from conftest import tol_ratio
from decompdiff_b200 import synthetic as syn
from oracle import restate

pytestmark = pytest.mark.gpu

DRIFT = [{'type': 'armsca_prox', 'min_d': 1.2, 'max_d': 1.9}, {'type': 'clash', 'sigma': 2, 'gamma': 4}]


def _check_forward_on_pockets(model, weights, cfg, kw, pockets, tag):
    out = model(**syn.forward_kwargs(kw, None))
    sub, rows = syn.select_pockets(kw, pockets)
    with torch.no_grad():
        ref = restate.forward(weights, cfg, **syn.forward_kwargs(sub, None))
    worst = {}
    for k, r in (('pred_ligand_pos', rows['ligand']), ('pred_ligand_v', rows['ligand']), ('pred_bond', rows['bond'])):
        worst[k] = tol_ratio(out[k].cpu()[r], ref[k])
    print(tag, 'pockets', pockets, 'max err / tol', {k: round(v, 4) for k, v in worst.items()})
    for k, v in worst.items():
        assert v <= 1.0, f'{tag}:{k} {v:.3f} x tol'
    return out


def test_forward_cfg2_headline_batch(model_cpu, weights, oracle_cfg):
    """bench.py's default workload: 64 pockets x (370 protein + 30 ligand atoms)."""
    kw = syn.make_batch(n_pockets=64, n_protein=370, arm_sizes=(8, 8), n_scaffold=14, seed=1000)
    _check_forward_on_pockets(model_cpu, weights, oracle_cfg, kw, [0, 7, 13, 22, 31, 40, 55, 63], 'cfg2 B=64')


@pytest.mark.parametrize('n_atoms', [100, 1000])
def test_forward_cfg5_sweep_ends(n_atoms, model_cpu, weights, oracle_cfg):
    """cfg 5: 32 pockets at both ends of the atom-count sweep (N = 100 and N = 1000 atoms per complex)."""
    kw = syn.make_batch(n_pockets=32, n_protein=n_atoms - 30, arm_sizes=(8, 8), n_scaffold=14, seed=700 + n_atoms)
    _check_forward_on_pockets(model_cpu, weights, oracle_cfg, kw, [0, 15, 31] if n_atoms == 1000 else [0, 9, 15, 22, 31],
                              f'cfg5 N={n_atoms}')


@pytest.mark.parametrize('arms,scaffold', [((10, 10), 20), ((13, 12), 25), ((16, 16), 32)], ids=['n40', 'n50', 'n64'])
def test_forward_large_ligands(arms, scaffold, model_cpu, weights, oracle_cfg):
    """Ligands of 40 / 50 / 64 atoms: bond-edge and triplet groups of more than 32 rows."""
    kw = syn.make_batch(n_pockets=4, n_protein=360, arm_sizes=arms, n_scaffold=scaffold, seed=sum(arms) + scaffold)
    _check_forward_on_pockets(model_cpu, weights, oracle_cfg, kw, [0, 3], f'n_lig={sum(arms) + scaffold}')


def test_forward_mixed_ligand_sizes_in_one_batch(model_cpu, weights, oracle_cfg):
    """One large ligand must not change how the other pockets of the batch are computed."""
    g = torch.Generator().manual_seed(5)
    pockets = [syn.make_pocket(g, 300, (8, 8), 14), syn.make_pocket(g, 300, (12, 12), 20), syn.make_pocket(g, 200, (3,), 2)]
    kw = syn.collate_pockets(pockets)
    _check_forward_on_pockets(model_cpu, weights, oracle_cfg, kw, [0, 1, 2], 'mixed ligand sizes')


@pytest.mark.parametrize('t', [999, 500, 0])
def test_teacher_forced_guided_step_cfg3_batch(t, model_cpu, weights, oracle_cfg):
    """cfg 3: one reverse step (network + armsca_prox / clash drift + posterior) of the 64-pocket guided batch against the
    oracle on 6 of its pockets, with injected noise.  The armsca energy is divided by the batch size (guidance_funcs.py:78),
    so the oracle on the slice is told the full batch size."""
    kw = syn.make_batch(n_pockets=64, n_protein=370, arm_sizes=(8, 8), n_scaffold=14, seed=1003, n_full_extra=2000)
    n, Eb = kw['init_ligand_pos'].size(0), kw['init_ligand_fc_bond_type'].numel()
    noise = syn.step_noise(n, Eb, 1, seed=77)
    run = model_cpu.begin_sampling(**kw, num_steps=1, center_pos_mode='protein', energy_drift_opt=DRIFT)
    run.eb.set_time(t)
    run.advance(1, noise=noise)
    got = run.finish(traj_on_device=True)
    pockets = [0, 11, 29, 30, 47, 63]
    sub, rows = syn.select_pockets(kw, pockets)
    sub_noise = [{'u_atom': noise[0]['u_atom'][rows['ligand']], 'u_bond': noise[0]['u_bond'][rows['bond']],
                  'eps_pos': noise[0]['eps_pos'][rows['ligand']]}]
    want = restate.sample_diffusion(weights, oracle_cfg, **sub, num_steps=1, center_pos_mode='protein', energy_drift_opt=DRIFT,
                                    noise=sub_noise, first_t=t, num_graphs_div=64)
    r_pos = tol_ratio(got['pos'].cpu()[rows['ligand']], want['pos'])
    flips_v = int((got['v'].cpu()[rows['ligand']] != want['v']).sum())
    flips_b = int((got['bond'].cpu()[rows['bond']] != want['bond']).sum())
    r_vt = tol_ratio(got['vt_traj'][0].cpu()[rows['ligand']], want['vt_traj'][0])
    r_bt = tol_ratio(got['bt_traj'][0].cpu()[rows['bond']], want['bt_traj'][0])
    print(f'cfg3 B=64 t={t}: pos {r_pos:.3f} x tol, log p(v) {r_vt:.3f}, log p(b) {r_bt:.3f}, flips v {flips_v} b {flips_b}')
    assert r_pos <= 1.0 and r_vt <= 1.0 and r_bt <= 1.0
    assert flips_v == 0 and flips_b <= 2      # a Gumbel near-tie may flip a bond sample


def test_wide_weights_stress_the_tf32_split(model_cpu, oracle_cfg):
    """Every tensor-core product is a truncating hi/lo TF32 split (attn_tc.cuh tf32_split); default-initialised weights have a
    narrow dynamic range.  Here the Linear weights are scaled by 4 and the LayerNorm gains spread over [0.25, 4], which widens the
    range of every hidden activation; the outputs must stay inside the same tolerance."""
    import decompdiff_b200 as ddb
    m = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
    sd = syn.synthetic_state_dict(m, seed=3)
    g = torch.Generator().manual_seed(8)
    for k in sd:
        if k.startswith('refine_net') and k.endswith('net.0.weight'):
            sd[k] = sd[k] * 4.0
        elif k.startswith('refine_net') and k.endswith('net.1.weight'):
            sd[k] = sd[k] * torch.exp2(4 * torch.rand(sd[k].shape, generator=g) - 2)
    m.load_state_dict(sd)
    kw = syn.make_batch(n_pockets=3, n_protein=200, arm_sizes=(6, 5), n_scaffold=9, seed=44)
    fk = syn.forward_kwargs(kw, None)
    out = m.eval()(**fk)
    with torch.no_grad():
        ref = restate.forward(sd, oracle_cfg, **fk)
    for k in ('pred_ligand_pos', 'pred_ligand_v', 'pred_bond'):
        r = tol_ratio(out[k], ref[k])
        print('wide weights', k, 'max err / tol', round(r, 4), 'max |ref|', float(ref[k].abs().max()))
        assert r <= 1.0, k
