"""Host-side rows of the hot path (a1 sampling driver, a16 prior set-up / transforms) against the
UNMODIFIED reference functions.  The fixtures (tests/golden/driver_*.pt) were produced by oracle/make_golden_driver.py, which
runs the reference's own `sample_diffusion_ligand_decomp`, `utils.prior` and `utils.transforms` on seeded synthetic pockets
with a recording stub model; here the product's functions run on the same seeds and must hand the model bit-identical
tensors and un-batch bit-identical results.  Integer / index work: exact.  Float work on the host (float32 torch, same op
order): exact as well."""
import logging

import numpy as np
import pytest
import torch

from conftest import load_golden
from decompdiff_b200 import prior, sampling, synthetic as syn, transforms as trans
from oracle.make_golden_driver import (ATOM_PRIOR, BATCH_SIZE, BOND_PRIOR, DRIVER_CASES, NUM_SAMPLES, NUM_STEPS, SEED, StubModel,
                                       build_case, natoms_configs, pack_results, stat_models)

DRIFT = [{'type': 'armsca_prox', 'min_d': 1.2, 'max_d': 1.9}, {'type': 'clash', 'sigma': 2, 'gamma': 4}]


def run_driver(spec, model, device='cpu', **over):
    data, init_transform, full_pos = build_case(spec, trans, prior, trans.Compose)
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    kw = dict(init_transform=init_transform, num_samples=NUM_SAMPLES, batch_size=BATCH_SIZE, device=device,
              prior_mode=spec['prior_mode'], num_steps=NUM_STEPS, center_pos_mode='protein', num_atoms_mode=spec['num_atoms_mode'],
              atom_prior_probs=ATOM_PRIOR if spec['type_priors'] else None, bond_prior_probs=BOND_PRIOR if spec['type_priors'] else None,
              atom_enc_mode='basic', bond_fc_mode='fc', energy_drift_opt=DRIFT, full_protein_pos=full_pos,
              arms_natoms_config=natoms_configs(spec)[0], scaffold_natoms_config=natoms_configs(spec)[1],
              natoms_config=stat_models(spec))
    kw.update(over)
    return sampling.sample_diffusion_ligand_decomp(model, data, **kw)


def assert_same(a, b, where):
    if torch.is_tensor(b):
        assert torch.is_tensor(a), where
        assert a.dtype == b.dtype and a.shape == b.shape, f'{where}: {a.dtype}{tuple(a.shape)} vs {b.dtype}{tuple(b.shape)}'
        assert torch.equal(a, b), f'{where}: max |d| = {float((a.double() - b.double()).abs().max())}'
    else:
        assert a == b, where


@pytest.mark.parametrize('name', list(DRIVER_CASES))
def test_driver_matches_reference(name):
    """Every tensor handed to model.sample_diffusion and every un-batched result equals the reference's, bit for bit."""
    gold = load_golden(f'driver_{name}')
    model = StubModel()
    got = pack_results(model, run_driver(DRIVER_CASES[name], model))
    assert len(got['calls']) == len(gold['calls']) == 3          # mini-batches of 2 + 2 + 1
    for ci, (c_got, c_ref) in enumerate(zip(got['calls'], gold['calls'])):
        assert set(c_got) == set(c_ref), f'call {ci}: keyword arguments differ: {set(c_got) ^ set(c_ref)}'
        for k in c_ref:
            assert_same(c_got[k], c_ref[k], f'{name} call {ci} {k}')
    assert len(got['results']) == len(gold['results']) == NUM_SAMPLES
    for ri, (r_got, r_ref) in enumerate(zip(got['results'], gold['results'])):
        assert set(r_got) == set(r_ref)
        for k in r_ref:
            assert_same(r_got[k], r_ref[k], f'{name} result {ri} {k}')
        assert r_got['pred_pos'].dtype == torch.float64 and r_got['pred_pos_traj'].dtype == torch.float64


def test_driver_result_schema_and_edges():
    spec = DRIVER_CASES['ref_prior']
    res = run_driver(spec, StubModel(), logger=logging.getLogger('t'))
    assert set(res[0]) == {'mol', 'smiles', 'pred_pos', 'pred_v', 'pred_pos_traj', 'pred_v_traj', 'decomp_mask',
                           'pred_bond_index', 'pred_bond_type'}
    assert res[0]['mol'] is None and res[0]['smiles'] == ''
    n = len(res[0]['decomp_mask'])
    assert res[0]['pred_pos'].shape == (n, 3) and res[0]['pred_pos_traj'].shape == (NUM_STEPS, n, 3)
    assert np.asarray(res[0]['pred_bond_index']).shape == (2, n * (n - 1))
    assert np.asarray(res[0]['pred_bond_index']).min() == 0 and np.asarray(res[0]['pred_bond_index']).max() == n - 1
    # a reconstruction hook sees atomic numbers of the 'basic' vocabulary and its result lands in the dict
    seen = []
    res = run_driver(spec, StubModel(), reconstruct_fn=lambda pos, z, arom, bi, bt: (seen.append((z, arom)) or ('MOL', 'CC')))
    assert res[0]['mol'] == 'MOL' and res[0]['smiles'] == 'CC' and seen[0][1] is None
    assert set(seen[0][0]) <= {1, 6, 7, 8, 9, 15, 16, 17}
    # errors: unknown prior mode / atom-count mode (ValueError as the reference)
    with pytest.raises(ValueError):
        run_driver(spec, StubModel(), prior_mode='nope')
    with pytest.raises(ValueError):
        run_driver(DRIVER_CASES['subpocket_ref'], StubModel(), num_atoms_mode='nope')
    # one sample, batch larger than the request
    assert len(run_driver(spec, StubModel(), num_samples=1, batch_size=4)) == 1


def test_result_file_round_trip(tmp_path):
    """result.pt hand-off (scripts/sample_diffusion_decomp.py:609-619): list of dicts + ligand_filename, loadable with torch.load."""
    res = run_driver(DRIVER_CASES['ref_prior'], StubModel())
    path = str(tmp_path / 'result.pt')
    saved = sampling.save_results(res, path, ligand_filename='pocket/lig.sdf')
    loaded = torch.load(path, weights_only=False)
    assert len(loaded) == len(res) == NUM_SAMPLES and loaded[0]['ligand_filename'] == 'pocket/lig.sdf'
    assert set(loaded[0]) == set(saved[0]) == set(res[0]) | {'ligand_filename'}
    for a, b in zip(loaded, res):
        assert np.array_equal(a['pred_pos'], b['pred_pos']) and a['pred_pos'].dtype == np.float64
        assert np.array_equal(a['pred_bond_type'], b['pred_bond_type']) and a['pred_bond_index'] == b['pred_bond_index']


def test_priors_and_transforms_match_reference():
    gold = load_golden('driver_transforms')
    d = syn.make_raw_pocket(seed=41, arm_sizes=(3, 1, 4), n_scaffold=4)
    prior.compute_golden_prior_from_data(d)
    for part, entries in (('arms', d.arms_prior), ('scaffold', d.scaffold_prior)):
        assert len(entries) == len(gold['golden_prior'][part])
        for got, ref in zip(entries, gold['golden_prior'][part]):
            assert got[0] == ref[0]
            for g, r in zip(got[1:], ref[1:]):
                g = torch.as_tensor(g)
                assert g.dtype == r.dtype and torch.equal(g, r)
    assert torch.equal(d.pocket_prior_masks, gold['golden_prior']['pocket_prior_masks'])
    prior.apply_std_coef(d, 1.5)
    prior.apply_num_atoms_change(d, -2)
    for part, entries in (('arms', d.arms_prior), ('scaffold', d.scaffold_prior)):
        for got, ref in zip(entries, gold['rescaled_prior'][part]):
            assert got[0] == ref[0] and torch.equal(torch.as_tensor(got[2]), ref[1])

    d = trans.FeaturizeProteinAtom()(syn.make_raw_pocket(seed=42, arm_sizes=(3, 2), n_scaffold=4))
    assert d.protein_atom_feature.shape[1] == trans.FeaturizeProteinAtom().protein_feature_dim == 27
    prior.compute_golden_prior_from_data(d)
    d = trans.ComputeLigandAtomNoiseDist('ref_prior')(d)
    ind = trans.AddDecompIndicator(max_num_arms=10, global_prior_index=8, add_ord_feat=True)
    assert ind.protein_feature_dim == ind.ligand_feature_dim == 13
    d = ind(d)
    for k, ref in gold['ord_feat'].items():
        assert_same(getattr(d, k), ref, k)
    for mode in ('decomp_fc', 'scaffold_fc'):
        assert_same(trans.FeaturizeLigandBond(mode=mode)(d).ligand_fc_bond_index, gold[f'bond_{mode}'], mode)
    d.ligand_bond_index = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    d.ligand_bond_type = torch.tensor([1, 1, 2, 2])
    assert_same(trans.FeaturizeLigandBond(mode='fc', set_bond_type=True)(d).ligand_fc_bond_type, gold['bond_fc_typed'], 'typed')
    with pytest.raises(ValueError):
        trans.FeaturizeLigandBond(mode='nope')(d)
    for m, n in (('basic', 8), ('add_aromatic', 13), ('full', 23)):
        assert trans.get_atomic_number_from_index(torch.arange(n), m) == gold['atomic_numbers'][m]
        assert trans.is_aromatic_from_index(torch.arange(n), m) == gold['aromatic'][m]
        assert trans.FeaturizeLigandAtom(m, prior_types=False).ligand_feature_dim == n
    with pytest.raises(ValueError):
        trans.get_atomic_number_from_index(torch.arange(3), 'nope')
    f = trans.FeaturizeLigandAtom('basic', prior_types=True)
    assert np.array_equal(f.atom_types_prob, ATOM_PRIOR) and np.array_equal(f.bond_types_prob, BOND_PRIOR)


def test_prior_std_rules():
    """min-std clamp 0.6, single-atom parts, missing scaffold row, scalar beta variances (utils/transforms.py:195-245)."""
    d = syn.make_raw_pocket(seed=43, arm_sizes=(1, 5), n_scaffold=0)
    prior.compute_golden_prior_from_data(d)
    d = trans.ComputeLigandAtomNoiseDist('ref_prior')(d)
    assert d.ligand_decomp_centers.shape == (3, 3) and d.ligand_decomp_stds.shape == (3, 3)
    assert torch.all(d.ligand_decomp_stds[0] == 0.6) and torch.all(d.ligand_decomp_stds[2] == 0.6)
    assert torch.allclose(d.ligand_decomp_centers[2], d.protein_pos.mean(0))
    assert d.ligand_decomp_num_atoms.tolist() == [1, 5, 0]
    d = syn.make_raw_pocket(seed=44)
    prior.substitute_golden_prior_with_given_prior(d, syn.beta_prior_dict(3, scalar_scaffold_cov=True))
    assert d.pocket_atom_masks.shape == (2, 60) and d.pocket_atom_masks.dtype == torch.bool
    d = trans.ComputeLigandAtomNoiseDist('beta_prior')(d)
    assert bool((d.ligand_decomp_stds >= 0.6).all()) and d.ligand_decomp_stds.dtype == torch.float32
    with pytest.raises(AssertionError):
        trans.ComputeLigandAtomNoiseDist('nope')


@pytest.mark.gpu
def test_driver_end_to_end_on_gpu(model_cpu):
    """The real model behind the driver on cuda:0: same molecules as calling sample_diffusion on the driver's own batch."""
    spec = DRIVER_CASES['ref_prior']
    res = run_driver(spec, model_cpu, device='cuda:0', num_samples=3, batch_size=2, num_steps=4)
    assert len(res) == 3
    for r in res:
        n = len(r['decomp_mask'])
        assert r['pred_pos'].shape == (n, 3) and r['pred_pos'].dtype == np.float64 and np.isfinite(r['pred_pos']).all()
        assert r['pred_pos_traj'].shape == (4, n, 3) and r['pred_v_traj'].shape == (4, n)
        assert np.array_equal(r['pred_pos_traj'][-1], r['pred_pos']) and np.array_equal(r['pred_v_traj'][-1], r['pred_v'])
        assert r['pred_bond_type'].shape == (n * (n - 1),) and 0 <= r['pred_bond_type'].min() and r['pred_bond_type'].max() < 5


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['ref_prior', 'beta_prior_v2', 'beta_prior_old'])
def test_driver_values_on_gpu_against_the_oracle(name, model_cpu, weights, oracle_cfg):
    """VALUES of the driver's results on cuda:0, not only shapes: the driver runs with the CUDA model behind a thin wrapper that
    records the tensors of every mini-batch call and injects seeded noise; the CPU oracle then runs `sample_diffusion` on exactly
    those tensors with the same noise, and the driver's un-batched molecules / trajectories must be the oracle's (discrete samples
    equal, positions within tolerance).  Covers the replicated collate (fixed atom counts) and the per-sample collate ('old')."""
    from conftest import tol_ratio
    from oracle import restate

    class Recorder:
        num_classes, num_bond_classes, bond_diffusion = model_cpu.num_classes, model_cpu.num_bond_classes, True

        def __init__(self):
            self.calls = []

        def sample_diffusion(self, **kw):
            n, eb = kw['init_ligand_pos'].size(0), kw['init_ligand_fc_bond_type'].numel()
            noise = syn.step_noise(n, eb, kw['num_steps'], seed=100 + len(self.calls))
            self.calls.append(({k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in kw.items()}, noise))
            assert kw['protein_pos'].is_cuda and kw['ligand_fc_bond_index'].is_cuda      # the batch was assembled on the device
            return model_cpu.sample_diffusion(**kw, noise=noise)

    rec = Recorder()
    res = run_driver(DRIVER_CASES[name], rec, device='cuda:0', num_samples=3, batch_size=2, num_steps=3)
    assert len(res) == 3 and len(rec.calls) == 2
    k = 0
    for kw, noise in rec.calls:
        want = restate.sample_diffusion(weights, oracle_cfg, **kw, noise=noise)
        atoms = torch.bincount(kw['batch_ligand']).tolist()
        bonds = torch.bincount(kw['batch_ligand_bond']).tolist()
        pos, v, bond = want['pos'].split(atoms), want['v'].split(atoms), want['bond'].split(bonds)
        traj = torch.stack(want['pos_traj']).split(atoms, dim=1)
        for i in range(len(atoms)):
            r = res[k]
            assert tol_ratio(torch.from_numpy(r['pred_pos']).float(), pos[i]) <= 1.0
            assert np.array_equal(r['pred_v'], v[i].numpy()) and np.array_equal(r['pred_bond_type'], bond[i].numpy())
            assert tol_ratio(torch.from_numpy(r['pred_pos_traj']).float(), traj[i]) <= 1.0
            k += 1
