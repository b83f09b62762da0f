"""TEST INFRASTRUCTURE ONLY - golden vectors for the host-side rows (a1 driver, a16 prior set-up).

Run in the build container (needs /root/reference):   python -m oracle.make_golden_driver
The UNMODIFIED reference functions are imported through `oracle/ref_shims.py`:
  * `utils.prior.compute_golden_prior_from_data / substitute_golden_prior_with_given_prior / apply_*`
  * `utils.transforms.{FeaturizeProteinAtom, ComputeLigandAtomNoiseDist, AddDecompIndicator, FeaturizeLigandBond}`
  * `scripts/sample_diffusion_decomp.py: sample_diffusion_ligand_decomp` driven with a recording stub model
    (`StubModel`, shared with the tests), so the fixture holds (1) every tensor the reference hands to
    `model.sample_diffusion` and (2) what it un-batches from the stub's deterministic outputs.
Inputs are regenerated from seeds (`decompdiff_b200.synthetic.make_raw_pocket / beta_prior_dict`); only the reference's
outputs are stored in tests/golden/driver_*.pt.
"""
from __future__ import annotations

import importlib.util
import logging
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from decompdiff_b200 import synthetic as syn  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, 'tests', 'golden')
ATOM_PRIOR = np.array([0., 0.6716, 0.1174, 0.1689, 0.01315, 0.01117, 0.01128, 0.00647])
BOND_PRIOR = np.array([0.9170, 0.0433, 0.00687, 0.000173, 0.03266])

# name -> driver arguments; 5 samples in mini-batches of 2 (2 + 2 + 1) exercise the ragged last batch
DRIVER_CASES = {
    'ref_prior': dict(pocket=dict(seed=31), prior_mode='ref_prior', num_atoms_mode='ref', type_priors=False),
    'ref_prior_typepriors': dict(pocket=dict(seed=32, arm_sizes=(1, 4)), prior_mode='ref_prior', num_atoms_mode='ref', type_priors=True),
    'ref_prior_noscaffold': dict(pocket=dict(seed=33, arm_sizes=(3, 2, 2), n_scaffold=0), prior_mode='ref_prior', num_atoms_mode='ref',
                                 type_priors=False),
    'beta_prior_v2': dict(pocket=dict(seed=34), prior_mode='beta_prior', num_atoms_mode='v2', type_priors=False, beta_seed=5),
    'beta_prior_old': dict(pocket=dict(seed=35), prior_mode='beta_prior', num_atoms_mode='old', type_priors=False, beta_seed=6,
                           beta_matrix_cov=True),
    'subpocket_ref': dict(pocket=dict(seed=36), prior_mode='subpocket', num_atoms_mode='ref', type_priors=False),
    'subpocket_ref_large': dict(pocket=dict(seed=37), prior_mode='subpocket', num_atoms_mode='ref_large', type_priors=False),
    # atom counts drawn from binned distributions: the bin comes from the reference's BUILT-IN bounds whatever the passed
    # dictionaries say (utils/evaluation/atom_num.py:20-35); one dictionary, one None (built-in distributions)
    # atom counts / stds from regressors (utils/prior.py:162-208); 3 arms: the reference's zip over (distances, stds) only
    # broadcasts for 1 or 3 arms
    'beta_prior_stat': dict(pocket=dict(seed=39, arm_sizes=(2, 3, 2), n_scaffold=4), prior_mode='beta_prior', num_atoms_mode='stat',
                            type_priors=False, beta_seed=8, stat_seed=4),
    'subpocket_prior': dict(pocket=dict(seed=38), prior_mode='subpocket', num_atoms_mode='prior', type_priors=False, natoms_seed=3),
}
NUM_SAMPLES, BATCH_SIZE, NUM_STEPS, SEED = 5, 2, 3, 2021


def natoms_configs(spec):
    """(arms_natoms_config, scaffold_natoms_config) of a case: a synthetic stand-in for the shipped arm_num_config.pkl (its own
    'bounds' differ from the built-in ones and must be ignored) and None for the scaffold."""
    if 'natoms_seed' not in spec:
        return None, None
    rng = np.random.RandomState(spec['natoms_seed'])
    bins = []
    for _ in range(10):
        counts = list(rng.choice(np.arange(2, 9), size=4, replace=False))
        p = rng.rand(4) + 0.1
        bins.append(([int(c) for c in counts], list(p / p.sum())))
    return {'bounds': list(np.linspace(5.0, 9.0, 9)), 'bins': bins}, None


class LinearRegressor:
    """scikit-learn style stand-in for the pickled regressors of `num_atoms_mode='stat'`: predict(X) = X w + b."""

    def __init__(self, w, b):
        self.w, self.b = np.asarray(w, dtype=np.float64), float(b)

    def predict(self, X):
        return np.asarray(X, dtype=np.float64) @ self.w + self.b


def stat_models(spec):
    if 'stat_seed' not in spec:
        return None
    rng = np.random.RandomState(spec['stat_seed'])
    return {'arm_model': LinearRegressor(rng.rand(50) * 0.004, 3.0), 'armstd_model': LinearRegressor([0.08], 0.5),
            'sca_model': LinearRegressor(np.append(rng.rand(50) * 0.004, 0.05), 3.5), 'scastd_model': LinearRegressor([0.06], 0.7)}


class StubModel:
    """Stands in for DecompScorePosNet3D: records the keyword arguments of every `sample_diffusion` call and returns
    deterministic outputs derived from them (so the driver's un-batching can be checked without a network)."""
    num_classes, num_bond_classes, bond_diffusion = 8, 5, True

    def __init__(self):
        self.calls = []

    def sample_diffusion(self, **kw):
        self.calls.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items()})
        pos, v, b = kw['init_ligand_pos'], kw['init_ligand_v'], kw['init_ligand_fc_bond_type']
        n, eb, T = pos.size(0), b.numel(), kw['num_steps']
        ramp_v = torch.arange(n * 8, dtype=torch.float32).reshape(n, 8) / 7.0
        ramp_b = torch.arange(eb * 5, dtype=torch.float32).reshape(eb, 5) / 3.0
        return {
            'pos': pos * 0.5 + 1.0, 'v': (v + 1) % 8, 'bond': (b + 2) % 5,
            'pos_traj': [pos + float(s) for s in range(T)], 'v_traj': [(v + s) % 8 for s in range(T)],
            'v0_traj': [ramp_v + s for s in range(T)], 'vt_traj': [ramp_v - s for s in range(T)],
            'bond_traj': [(b + s) % 5 for s in range(T)], 'bt_traj': [ramp_b * (s + 1) for s in range(T)],
        }


def build_case(spec, trans, prior, Compose):
    """The pocket, its prior and the transform pipeline of scripts/sample_diffusion_decomp.py:509-590 for one case,
    built from the modules passed in (the reference's or the product's - same call sequence)."""
    data = syn.make_raw_pocket(**spec['pocket'])
    if spec['prior_mode'] == 'beta_prior':
        p = spec['pocket']
        bp = syn.beta_prior_dict(spec['beta_seed'], p.get('arm_sizes', (3, 4)), p.get('n_scaffold', 5),
                                 scalar_scaffold_cov=not spec.get('beta_matrix_cov', False))
        prior.substitute_golden_prior_with_given_prior(data, bp)
    elif spec['prior_mode'] == 'ref_prior':
        prior.compute_golden_prior_from_data(data)
    data = trans.FeaturizeProteinAtom()(data)
    init_transform = Compose([
        trans.ComputeLigandAtomNoiseDist(version=spec['prior_mode']),
        trans.AddDecompIndicator(max_num_arms=10, global_prior_index=8, add_ord_feat=False),
        trans.FeaturizeLigandBond(mode='fc', set_bond_type=False),
    ])
    full_protein_pos = torch.cat([data.protein_pos, torch.randn(20, 3, generator=torch.Generator().manual_seed(9)) * 20.0], 0)
    return data, init_transform, full_protein_pos


def pack_results(model, results):
    calls = [{k: v for k, v in c.items() if torch.is_tensor(v) or isinstance(v, (int, str, type(None), list))} for c in model.calls]
    out = []
    for r in results:
        out.append({'pred_pos': torch.from_numpy(np.asarray(r['pred_pos'])), 'pred_v': torch.from_numpy(np.asarray(r['pred_v'])),
                    'pred_pos_traj': torch.from_numpy(np.asarray(r['pred_pos_traj'])),
                    'pred_v_traj': torch.from_numpy(np.asarray(r['pred_v_traj'])),
                    'decomp_mask': list(r['decomp_mask']), 'pred_bond_index': r['pred_bond_index'],
                    'pred_bond_type': torch.from_numpy(np.asarray(r['pred_bond_type']))})
    return {'calls': calls, 'results': out}


def load_reference_driver():
    root = ref_shims.load_reference()
    spec = importlib.util.spec_from_file_location('ref_sample_diffusion_decomp', os.path.join(root, 'scripts', 'sample_diffusion_decomp.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.logger = logging.getLogger('ref_driver')
    mod.args = types.SimpleNamespace(recon_with_bond=True)
    return mod


def stat_path(spec):
    """The reference opens `natoms_config` as a pickle file (:72-74)."""
    if 'stat_seed' not in spec:
        return None
    import pickle
    import tempfile
    path = os.path.join(tempfile.gettempdir(), f'ddb_stat_models_{spec["stat_seed"]}.pkl')
    with open(path, 'wb') as f:
        pickle.dump(stat_models(spec), f)
    return path


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    drv = load_reference_driver()
    import utils.prior as ref_prior
    import utils.transforms as ref_trans
    from torch_geometric.transforms import Compose

    only = set(sys.argv[1:])
    for name, spec in DRIVER_CASES.items():
        if only and name not in only:
            continue
        data, init_transform, full_pos = build_case(spec, ref_trans, ref_prior, Compose)
        drv.full_protein_pos = full_pos
        model = StubModel()
        torch.manual_seed(SEED)
        np.random.seed(SEED)
        res = drv.sample_diffusion_ligand_decomp(
            model, data, init_transform=init_transform, num_samples=NUM_SAMPLES, batch_size=BATCH_SIZE, device='cpu',
            prior_mode=spec['prior_mode'], num_steps=NUM_STEPS, center_pos_mode='protein', num_atoms_mode=spec['num_atoms_mode'],
            atom_prior_probs=ATOM_PRIOR if spec['type_priors'] else None, bond_prior_probs=BOND_PRIOR if spec['type_priors'] else None,
            atom_enc_mode='basic', bond_fc_mode='fc', arms_natoms_config=natoms_configs(spec)[0],
            scaffold_natoms_config=natoms_configs(spec)[1], natoms_config=stat_path(spec),
            energy_drift_opt=[{'type': 'armsca_prox', 'min_d': 1.2, 'max_d': 1.9}, {'type': 'clash', 'sigma': 2, 'gamma': 4}])
        torch.save(pack_results(model, res), os.path.join(GOLDEN_DIR, f'driver_{name}.pt'))
        print(name, len(model.calls), 'calls,', len(res), 'results, atoms', [len(r['decomp_mask']) for r in res])

    if only:
        return
    # transforms / priors not reached by the driver cases
    extra = {}
    d = syn.make_raw_pocket(seed=41, arm_sizes=(3, 1, 4), n_scaffold=4)
    ref_prior.compute_golden_prior_from_data(d)
    snap = lambda entries: [(n,) + tuple(torch.as_tensor(t).clone() for t in rest) for n, *rest in entries]   # apply_* mutate in place
    extra['golden_prior'] = {'arms': snap(d.arms_prior), 'scaffold': snap(d.scaffold_prior),
                             'pocket_prior_masks': d.pocket_prior_masks.clone()}
    ref_prior.apply_std_coef(d, 1.5)
    ref_prior.apply_num_atoms_change(d, -2)
    extra['rescaled_prior'] = {'arms': [(n, torch.as_tensor(cov).clone()) for n, _, cov, _, _ in d.arms_prior],
                               'scaffold': [(n, torch.as_tensor(cov).clone()) for n, _, cov, _, _ in d.scaffold_prior]}
    d = ref_trans.FeaturizeProteinAtom()(syn.make_raw_pocket(seed=42, arm_sizes=(3, 2), n_scaffold=4))
    ref_prior.compute_golden_prior_from_data(d)
    d = ref_trans.ComputeLigandAtomNoiseDist('ref_prior')(d)
    d = ref_trans.AddDecompIndicator(max_num_arms=10, global_prior_index=8, add_ord_feat=True)(d)
    extra['ord_feat'] = {'ligand_atom_aux_feature': d.ligand_atom_aux_feature, 'protein_atom_feature': d.protein_atom_feature,
                         'ligand_decomp_centers': d.ligand_decomp_centers, 'ligand_decomp_stds': d.ligand_decomp_stds}
    for mode in ('decomp_fc', 'scaffold_fc'):
        extra[f'bond_{mode}'] = ref_trans.FeaturizeLigandBond(mode=mode)(d).ligand_fc_bond_index.clone()
    d.ligand_bond_index = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    d.ligand_bond_type = torch.tensor([1, 1, 2, 2])
    extra['bond_fc_typed'] = ref_trans.FeaturizeLigandBond(mode='fc', set_bond_type=True)(d).ligand_fc_bond_type.clone()
    extra['atomic_numbers'] = {m: ref_trans.get_atomic_number_from_index(torch.arange(n), m)
                               for m, n in (('basic', 8), ('add_aromatic', 13), ('full', 23))}
    extra['aromatic'] = {m: ref_trans.is_aromatic_from_index(torch.arange(n), m) for m, n in (('basic', 8), ('add_aromatic', 13), ('full', 23))}
    torch.save(extra, os.path.join(GOLDEN_DIR, 'driver_transforms.pt'))
    print('extra', list(extra))


if __name__ == '__main__':
    main()
