"""Sampling driver of the decomposed-prior path.

`sample_diffusion_ligand_decomp` mirrors the function of the same name in
/root/reference/scripts/sample_diffusion_decomp.py:57-457: per mini-batch it draws the atom counts and
x_T = mu_k + eps * sigma_k for every sample (`subpocket` :80-147, `ref_prior` :149-201, `beta_prior` :203-295),
draws the initial bond / atom types (:186-194, :305-312), collates the samples (:314-316), calls
`model.sample_diffusion` with the reference's keyword arguments (:329-360) and un-batches molecules and
trajectories into per-sample float64 / int64 numpy arrays (:366-410).  The random draws are made with the same
torch calls in the same order, so a CPU run with the same seed hands `model.sample_diffusion` bit-identical
tensors (tests/test_driver.py checks that against the reference's own function).

Differences, all at the edges of the hot path:
* the reference reads module globals (`full_protein_pos`, `logger`, `args.recon_with_bond`); here they are the keyword
  arguments `full_protein_pos`, `logger`, `reconstruct_fn`;
* RDKit reconstruction (:416-455) is out of scope: `mol` is None and `smiles` '' unless a
  `reconstruct_fn(pred_pos, atomic_numbers, aromatic, bond_index, bond_type)` is supplied;
* `num_atoms_mode='stat'`: `natoms_config` may be the path of the pickle (as in the reference) or the already loaded dictionary
  of regressors (`prior.NumAtomsSampler`).
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import transforms as trans
from .prior import NumAtomsSampler
from .batch import FOLLOW_BATCH, Batch

COLLATE_EXCLUDE_KEYS = ('scaffold_prior', 'arms_prior')      # sample_diffusion_decomp.py:314


def log_sample_categorical(logits: torch.Tensor) -> torch.Tensor:
    """Gumbel-max draw with the reference's constants (models/transitions.py:78-84)."""
    uniform = torch.rand_like(logits)
    gumbel = -torch.log(-torch.log(uniform + 1e-30) + 1e-30)
    return (gumbel + logits).argmax(dim=-1)


def _draw_types(n: int, num_classes: int, prior_probs, device) -> torch.Tensor:
    """Initial categorical state: multinomial from the type prior, else uniform via Gumbel-max (:186-194, :305-312)."""
    if prior_probs is not None:
        return torch.multinomial(torch.from_numpy(prior_probs.astype(np.float32)), n, replacement=True).to(device)
    return log_sample_categorical(torch.zeros(n, num_classes).to(device))


def get_space_size(pocket_pos: np.ndarray) -> float:
    """Median of the 10 largest pairwise distances (utils/evaluation/atom_num.py:14-17)."""
    diff = pocket_pos[:, None, :] - pocket_pos[None, :, :]
    iu = np.triu_indices(len(pocket_pos), k=1)
    d = np.sort(np.sqrt((diff ** 2).sum(-1))[iu])[::-1]
    return float(np.median(d[:10]))


_ATOM_NUM_CONFIG = None


def atom_num_config() -> Dict:
    """The reference's built-in table `atom_num_config.CONFIG` (bin bounds + per-bin atom-count distribution), exported to
    `data/atom_num_config.json` by `oracle/make_atom_num_config.py`."""
    global _ATOM_NUM_CONFIG
    if _ATOM_NUM_CONFIG is None:
        import json
        import os
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'atom_num_config.json')) as f:
            raw = json.load(f)
        _ATOM_NUM_CONFIG = {'bounds': raw['bounds'], 'bins': [(c, p) for c, p in raw['bins']]}
    return _ATOM_NUM_CONFIG


def sample_atom_num(space_size: float, config_dict: Optional[Dict] = None) -> int:
    """Atom count from the binned empirical distribution (utils/evaluation/atom_num.py:20-35).  As in the reference the bin
    index ALWAYS comes from the built-in bounds (`_get_bin_idx` reads CONFIG['bounds'], :20-25); a passed dictionary (the
    arm / scaffold pickles) only replaces the per-bin distributions, and `None` falls back to the built-in ones."""
    bounds = atom_num_config()['bounds']
    idx = next((i for i, b in enumerate(bounds) if b > space_size), len(bounds))
    counts, probs = (atom_num_config() if config_dict is None else config_dict)['bins'][idx]
    return int(np.random.choice(counts, p=probs))


class _Plan:
    """Where the atoms of one sample start: per-arm centres / stds / counts and the scaffold's."""

    def __init__(self, arm_centers, arm_stds, arm_counts, sca_center, sca_std, sca_count):
        self.arm_centers, self.arm_stds, self.arm_counts = arm_centers, arm_stds, arm_counts
        self.sca_center, self.sca_std, self.sca_count = sca_center, sca_std, sca_count


def _draw_sample(plan: _Plan):
    """x_T and the arm id per atom (-1 = scaffold) of one sample: arms in order, scaffold last; one randn per part."""
    pos, mask = [], []
    for a, (mu, n) in enumerate(zip(plan.arm_centers, plan.arm_counts)):
        eps = torch.randn([n, 3])
        pos.append(mu + (eps if plan.arm_stds is None else eps * plan.arm_stds[a].unsqueeze(0)))
        mask += [a] * n
    eps = torch.randn([plan.sca_count, 3])
    pos.append(plan.sca_center + (eps if plan.sca_std is None else eps * plan.sca_std.unsqueeze(0)))
    mask += [-1] * plan.sca_count
    return torch.cat(pos, dim=0), mask


def _unbatch(seq, n_data: int, cum) -> List[np.ndarray]:
    """List over steps of flat arrays -> per sample (num_steps, n_i, ...) (unbatch_v_traj, :45-53)."""
    arrays = [x.cpu().numpy() for x in seq]
    return [np.stack([a[cum[k]:cum[k + 1]] for a in arrays]) for k in range(n_data)]


@torch.no_grad()
def sample_diffusion_ligand_decomp(
        model, data, init_transform, num_samples, batch_size=16, device='cuda:0', prior_mode='subpocket',
        num_steps=None, center_pos_mode='none', num_atoms_mode='prior',
        atom_prior_probs=None, bond_prior_probs=None,
        arms_natoms_config=None, scaffold_natoms_config=None, natoms_config=None,
        atom_enc_mode='add_aromatic', bond_fc_mode='fc', energy_drift_opt=None,
        full_protein_pos: Optional[torch.Tensor] = None, logger: Optional[logging.Logger] = None,
        reconstruct_fn: Optional[Callable] = None):
    """Sample `num_samples` molecules for the pocket `data`; returns the reference's list of result dicts
    (`pred_pos` (n,3) f64, `pred_v` (n,), `pred_pos_traj` (T,n,3), `pred_v_traj`, `decomp_mask`, `pred_bond_index`,
    `pred_bond_type`, `mol`, `smiles`)."""
    if prior_mode not in ('subpocket', 'ref_prior', 'beta_prior'):
        raise ValueError(prior_mode)
    natoms_sampler = None
    if num_atoms_mode == 'stat':      # :72-75
        if isinstance(natoms_config, (str, bytes)):
            import pickle
            with open(natoms_config, 'rb') as f:
                natoms_config = pickle.load(f)
        natoms_sampler = NumAtomsSampler(natoms_config)
    if full_protein_pos is None:
        full_protein_pos = data.protein_pos
    num_batch = int(np.ceil(num_samples / batch_size))
    results_per_sample: List[Dict] = []

    for i in range(num_batch):
        n_data = batch_size if i < num_batch - 1 else num_samples - batch_size * (num_batch - 1)

        # ---- where each part of the ligand starts (identical for every sample of the mini-batch unless counts are drawn)
        if prior_mode == 'subpocket':
            if num_atoms_mode == 'prior':
                arm_sizes = [get_space_size(data.protein_pos[m].detach().cpu().numpy()) for m in data.pocket_atom_masks]
                sca_size = get_space_size(data.protein_pos.detach().cpu().numpy())
            elif num_atoms_mode not in ('ref', 'ref_large'):
                raise ValueError(num_atoms_mode)
            arm_centers = [data.protein_pos[m].mean(0) for m in data.pocket_atom_masks]
            sca_center = data.protein_pos.mean(0)
            data.ligand_decomp_centers = torch.cat(arm_centers + [sca_center], dim=0)     # as the reference (:92)
            arm_stds = sca_std = None                                                       # unit variance

            def plan_for_sample():
                if num_atoms_mode == 'prior':
                    arms = [sample_atom_num(s, arms_natoms_config) for s in arm_sizes]
                    sca = sample_atom_num(sca_size, scaffold_natoms_config)
                else:
                    inc = int(np.ceil(10 / (data.num_arms + 2))) if num_atoms_mode == 'ref_large' else 0
                    arms = [int((data.ligand_atom_mask == a).sum()) + inc for a in range(data.num_arms)]
                    sca = int((data.ligand_atom_mask == -1).sum()) + 2 * inc
                return _Plan(arm_centers, None, arms, sca_center, None, sca)
        else:
            old = init_transform(data.clone())
            arm_centers = old.ligand_decomp_centers[:old.num_arms, :]
            sca_center = old.ligand_decomp_centers[-1, :]
            arm_stds = old.ligand_decomp_stds[:old.num_arms, :]
            sca_std = old.ligand_decomp_stds[-1, :]
            if prior_mode == 'ref_prior':
                arm_counts = [int(old.arms_prior[a][0]) for a in range(data.num_arms)]
                sca_count = int(old.scaffold_prior[0][0]) if len(old.scaffold_prior) == 1 else 0
            elif num_atoms_mode == 'v2':
                arm_counts = [int(data.arms_prior[a][0]) for a in range(data.num_arms)]
                sca_count = int(data.scaffold_prior[0][0]) if len(data.scaffold_prior) > 0 else 0
            elif num_atoms_mode not in ('old', 'stat'):
                raise ValueError(num_atoms_mode)

            def plan_for_sample():
                if num_atoms_mode == 'stat':      # counts and stds predicted from the pocket, drawn per sample (:221-231)
                    natoms, stds = natoms_sampler.sample_arm_natoms(arm_centers, data.protein_pos)
                    if len(data.scaffold_prior) > 0:
                        center = [p[1] for p in data.scaffold_prior][0]
                        sca_n, sca_s = natoms_sampler.sample_sca_natoms(center, arm_centers, stds, data.protein_pos)
                    else:
                        center, sca_n, sca_s = data.protein_pos.mean(0), 0, torch.tensor([0.])
                    return _Plan(arm_centers, stds, [int(n) for n in natoms], center, sca_s, int(sca_n))
                if prior_mode == 'beta_prior' and num_atoms_mode == 'old':
                    # atom count ~ U{lower..upper} from a linear fit on the prior std (:238-246, :257-263)
                    m, b = 12.41, -4.98
                    return _OldCountPlan(arm_centers, arm_stds, sca_center, sca_std, m, b)
                return _Plan(arm_centers, arm_stds, arm_counts, sca_center, sca_std, sca_count)

        # ---- per sample: x_T, decomposition mask, transforms, initial bond types (draw order = the reference's)
        # With fixed atom counts every sample of the mini-batch has the same decomposition mask, so the (RNG-free) transforms
        # give the same result: they run ONCE, the per-sample loop only makes the reference's draws in the reference's order, and
        # the batch is assembled on the device from one copy of the shared tensors (`Batch.from_replicas`, bit-identical to the
        # per-sample collate).  Drawn counts ('prior', 'old', 'stat') keep the per-sample path.
        fixed_counts = (prior_mode == 'subpocket' and num_atoms_mode in ('ref', 'ref_large')) or \
                       (prior_mode == 'ref_prior') or (prior_mode == 'beta_prior' and num_atoms_mode == 'v2')
        samples, init_pos, ligand_num_atoms, decomp_ind, noise_stds, bond_types = [], [], [], [], [], []
        proto = None
        for _ in range(n_data):
            plan = plan_for_sample()
            pos, mask = plan.draw() if isinstance(plan, _OldCountPlan) else _draw_sample(plan)
            if num_atoms_mode == 'stat':      # the predicted stds replace the collated prior stds (:249-250, :264-265, :322-323)
                noise_stds += [plan.arm_stds[a, :].unsqueeze(0).expand(1, 3) for a in range(len(plan.arm_counts))]
                noise_stds.append(plan.sca_std.unsqueeze(0).expand(1, 3))
            if fixed_counts and proto is not None:
                new_data = proto
            else:
                new_data = data.clone()
                new_data.ligand_atom_mask = torch.tensor(mask, dtype=torch.long)
                new_data = init_transform(new_data)
                if fixed_counts:
                    proto = new_data
            if getattr(new_data, 'ligand_fc_bond_index', None) is not None:
                bt = _draw_types(new_data.ligand_fc_bond_index.size(1), model.num_bond_classes, bond_prior_probs, 'cpu')
                if fixed_counts:
                    bond_types.append(bt)
                else:
                    new_data.ligand_fc_bond_type = bt
            if not fixed_counts:
                samples.append(new_data)
            init_pos.append(pos)
            ligand_num_atoms.append(len(mask))
            decomp_ind.append(new_data.ligand_atom_mask.tolist())
        if logger is not None:
            logger.info(f'ligand_num_atoms={ligand_num_atoms}')

        # ---- collate + H2D (the boundary of the hot path, :303-321)
        init_ligand_pos = torch.cat(init_pos, dim=0).to(device)
        batch_ligand = torch.repeat_interleave(torch.arange(n_data), torch.tensor(ligand_num_atoms)).to(device)
        assert len(init_ligand_pos) == len(batch_ligand)
        init_ligand_v = _draw_types(len(batch_ligand), model.num_classes, atom_prior_probs, device)
        if fixed_counts:
            if bond_types:
                proto.ligand_fc_bond_type = bond_types[0]      # the key (and its follow_batch vector) exists as in the per-sample path
            batch = Batch.from_replicas(proto, n_data, exclude_keys=COLLATE_EXCLUDE_KEYS, follow_batch=FOLLOW_BATCH, device=device,
                                        overrides={'ligand_fc_bond_type': bond_types} if bond_types else None)
        else:
            batch = Batch.from_data_list(samples, exclude_keys=COLLATE_EXCLUDE_KEYS, follow_batch=FOLLOW_BATCH).to(device)
        batch_full_protein_pos = full_protein_pos.repeat(n_data, 1).to(device)
        full_batch_protein = torch.arange(n_data).repeat_interleave(len(full_protein_pos)).to(device)
        if num_atoms_mode == 'stat':
            batch.ligand_decomp_stds = torch.cat(noise_stds, dim=0).to(device)

        r = model.sample_diffusion(
            protein_pos=batch.protein_pos,
            protein_v=batch.protein_atom_feature.float(),
            batch_protein=batch.protein_element_batch,
            protein_group_idx=batch.protein_decomp_group_idx,
            init_ligand_pos=init_ligand_pos,
            init_ligand_v=init_ligand_v,
            ligand_v_aux=batch.ligand_atom_aux_feature.float(),
            batch_ligand=batch_ligand,
            ligand_group_idx=batch.ligand_decomp_group_idx,
            ligand_atom_mask=None,
            prior_centers=batch.ligand_decomp_centers,
            prior_stds=batch.ligand_decomp_stds,
            prior_num_atoms=batch.ligand_decomp_num_atoms,
            batch_prior=batch.ligand_decomp_centers_batch,
            prior_group_idx=batch.prior_group_idx,
            ligand_fc_bond_index=getattr(batch, 'ligand_fc_bond_index', None),
            init_ligand_fc_bond_type=getattr(batch, 'ligand_fc_bond_type', None),
            batch_ligand_bond=getattr(batch, 'ligand_fc_bond_type_batch', None),
            ligand_decomp_batch=batch.ligand_decomp_mask,
            ligand_decomp_index=batch.ligand_atom_mask,
            num_steps=num_steps,
            center_pos_mode=center_pos_mode,
            energy_drift_opt=energy_drift_opt,
            full_protein_pos=batch_full_protein_pos,
            full_batch_protein=full_batch_protein,
        )

        # ---- un-batch (:366-410): float64 positions, per-sample slices of every trajectory
        cum_atoms = np.cumsum([0] + ligand_num_atoms)
        pos_array = r['pos'].cpu().numpy().astype(np.float64)
        v_array = r['v'].cpu().numpy()
        pos_traj = [a.astype(np.float64) for a in _unbatch(r['pos_traj'], n_data, cum_atoms)]
        v_traj = _unbatch(r['v_traj'], n_data, cum_atoms)
        with_bonds = bool(getattr(model, 'bond_diffusion', False))
        if with_bonds:
            bond_array = r['bond'].cpu().numpy()
            bond_index_array = batch.ligand_fc_bond_index.cpu().numpy()
            num_bonds = torch.bincount(batch.ligand_fc_bond_type_batch, minlength=n_data).tolist()
            cum_bonds = np.cumsum([0] + num_bonds)
        for k in range(n_data):
            res = {
                'pred_pos': pos_array[cum_atoms[k]:cum_atoms[k + 1]],
                'pred_v': v_array[cum_atoms[k]:cum_atoms[k + 1]],
                'pred_pos_traj': pos_traj[k],
                'pred_v_traj': v_traj[k],
                'decomp_mask': decomp_ind[k],
            }
            if with_bonds:
                res['pred_bond_index'] = (bond_index_array[:, cum_bonds[k]:cum_bonds[k + 1]] - cum_atoms[k]).tolist()
                res['pred_bond_type'] = bond_array[cum_bonds[k]:cum_bonds[k + 1]]
            results_per_sample.append(res)

    # ---- hand-off to reconstruction (:416-455).  Without bond diffusion the reference's zip() over the empty bond lists
    # yields no results at all; that quirk is kept.
    results, n_recon, n_complete = [], 0, 0
    for idx, res in enumerate(results_per_sample):
        if 'pred_bond_type' not in res:
            continue
        mol, smiles = None, ''
        if reconstruct_fn is not None:
            try:
                mol, smiles = reconstruct_fn(res['pred_pos'], trans.get_atomic_number_from_index(res['pred_v'], atom_enc_mode),
                                             trans.is_aromatic_from_index(res['pred_v'], atom_enc_mode),
                                             res['pred_bond_index'], res['pred_bond_type'])
                n_recon += 1
            except Exception as exc:       # the reference catches its MolReconsError only; any failure is per-molecule here
                if logger is not None:
                    logger.warning(f'Reconstruct failed {idx}: {exc}')
                mol, smiles = None, ''
        if mol is not None and '.' not in smiles:
            n_complete += 1
        results.append({'mol': mol, 'smiles': smiles, **res})
    if logger is not None:
        logger.info(f'n_reconstruct: {n_recon} n_complete: {n_complete}')
    return results


class _OldCountPlan:
    """beta_prior with `num_atoms_mode='old'`: the count of each part is drawn right before its positions, so the
    torch.randint / torch.randn calls interleave exactly as in the reference (:238-263)."""

    def __init__(self, arm_centers, arm_stds, sca_center, sca_std, m, b):
        self.arm_centers, self.arm_stds, self.sca_center, self.sca_std, self.m, self.b = arm_centers, arm_stds, sca_center, sca_std, m, b

    def _count(self, std0, lower_round):
        lo = torch.clamp(lower_round((self.m - 2.0) * std0 + self.b), min=2).long()
        hi = torch.clamp(torch.ceil((self.m + 3.0) * std0 + self.b), min=2).long()
        return int(torch.randint(low=lo, high=hi + 1, size=(1,)))

    def draw(self):
        pos, mask = [], []
        for a in range(len(self.arm_centers)):
            n = self._count(self.arm_stds[a][0], torch.floor)
            pos.append(self.arm_centers[a] + torch.randn([n, 3]) * self.arm_stds[a, :].unsqueeze(0))
            mask += [a] * n
        n = self._count(self.sca_std[0], torch.ceil)
        pos.append(self.sca_center + torch.randn([n, 3]) * self.sca_std.unsqueeze(0))
        mask += [-1] * n
        return torch.cat(pos, dim=0), mask


def save_results(raw_results, path: str, ligand_filename: Optional[str] = None, extra: Optional[Dict] = None) -> List[Dict]:
    """Write the `result.pt` file of the reference driver (scripts/sample_diffusion_decomp.py:609-619): the list of per-sample
    dicts with `ligand_filename` added to each, saved with torch.save - the format `scripts/evaluate_mol_from_meta_full.py:45-63`
    reads (`pred_pos`, `pred_v`, `pred_bond_index`, `pred_bond_type`, `mol`, `smiles`, trajectories, `decomp_mask`)."""
    results = [{**r, 'ligand_filename': ligand_filename, **(extra or {})} for r in raw_results]
    torch.save(results, path)
    return results
