"""Synthetic pockets of the shapes BASELINE.json names (no dataset / checkpoint offline).

Generator parameters follow SURVEY.md section 8(d):
protein atoms `pos ~ N(0, 8^2 A)` (or a min-spacing-thinned ball with `dense=True`), 27+2 = 29
feature dims (6 element one-hot, 20 residue one-hot, backbone bit, 2-dim arm indicator), ligand of
`arm_sizes` arms + `n_scaffold` scaffold atoms, `mu_k ~ N(0, 3^2)`, `sigma_k ~ U(0.6, 1.6)`
broadcast to 3 dims, `x_T = mu_k + eps*sigma_k`, `v_T ~ U{0..7}`, directed fully-connected bond
index in the order of /root/reference/utils/transforms.py:331-337, `b_T ~ U{0..4}`.

The tensors are arranged exactly as the reference driver hands them to
`DecompScorePosNet3D.sample_diffusion` (/root/reference/scripts/sample_diffusion_decomp.py:329-360).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .batch import Batch, FOLLOW_BATCH, ProteinLigandData

MAX_NUM_ARMS = 10  # configs/training.yml: data.transform.max_num_arms
PROTEIN_ELEMENTS = torch.tensor([1, 6, 7, 8, 16, 34])  # utils/transforms.py:118


def fc_bond_index(n_atoms: int) -> torch.Tensor:
    """Directed fully connected ligand graph, dst-major (utils/transforms.py:331-337)."""
    full_dst = torch.repeat_interleave(torch.arange(n_atoms), n_atoms)
    full_src = torch.arange(n_atoms).repeat(n_atoms)
    keep = full_dst != full_src
    return torch.stack([full_src[keep], full_dst[keep]], dim=0)


def _protein_positions(gen: torch.Generator, n: int, dense: bool) -> torch.Tensor:
    if not dense:
        return torch.randn(n, 3, generator=gen) * 8.0
    # rejection-thinned ball (radius 12 A, >= 1.2 A spacing): closer to a real pocket's density
    pts: List[torch.Tensor] = []
    kept = torch.zeros(0, 3)
    while kept.size(0) < n:
        cand = torch.randn(4 * n, 3, generator=gen)
        cand = cand / cand.norm(dim=-1, keepdim=True) * 12.0 * torch.rand(4 * n, 1, generator=gen) ** (1 / 3)
        for c in cand:
            if kept.size(0) == 0 or (kept - c).norm(dim=-1).min() >= 1.2:
                kept = torch.cat([kept, c[None]], 0)
                if kept.size(0) == n:
                    break
    return kept


def make_pocket(gen: torch.Generator, n_protein: int = 370, arm_sizes: Sequence[int] = (8, 8),
                n_scaffold: int = 14, dense: bool = False) -> ProteinLigandData:
    """One synthetic complex with the attributes the sampling driver reads from a collated batch."""
    num_arms = len(arm_sizes)
    d = ProteinLigandData()
    d.protein_pos = _protein_positions(gen, n_protein, dense)
    elem_idx = torch.randint(0, 6, (n_protein,), generator=gen)
    d.protein_element = PROTEIN_ELEMENTS[elem_idx]
    aa = torch.randint(0, 20, (n_protein,), generator=gen)
    backbone = torch.randint(0, 2, (n_protein, 1), generator=gen)
    d.num_arms = num_arms
    d.num_scaffold = 1 if n_scaffold > 0 else 0
    d.max_decomp_group = MAX_NUM_ARMS + 1

    # decomposed prior (ref_prior/beta_prior produce the same tensors with other values)
    centers = torch.randn(num_arms + 1, 3, generator=gen) * 3.0
    stds = (0.6 + torch.rand(num_arms + 1, 1, generator=gen)).expand(-1, 3).clone()
    d.ligand_decomp_centers = centers
    d.ligand_decomp_stds = stds
    d.ligand_decomp_num_atoms = torch.tensor(list(arm_sizes) + [n_scaffold])

    mask: List[int] = []
    for a, s in enumerate(arm_sizes):
        mask += [a] * s
    mask += [-1] * n_scaffold
    d.ligand_atom_mask = torch.tensor(mask, dtype=torch.long)
    n_lig = len(mask)
    d.ligand_element = torch.full((n_lig,), 6, dtype=torch.long)

    # AddDecompIndicator (utils/transforms.py:284-319)
    d.prior_group_idx = torch.arange(num_arms + 1)
    d.ligand_decomp_mask = torch.where(d.ligand_atom_mask < 0, torch.tensor(num_arms), d.ligand_atom_mask)
    d.ligand_decomp_group_idx = d.ligand_decomp_mask.clone()
    d.ligand_atom_aux_feature = F.one_hot((d.ligand_atom_mask >= 0).long(), num_classes=2)
    # protein arm indicator: atoms within 10 A of an arm centre (utils/prior.py:62-64 for beta_prior)
    near = torch.zeros(n_protein, dtype=torch.bool)
    for a in range(num_arms):
        near |= (d.protein_pos - centers[a]).norm(dim=-1) < 10.0
    protein_arm_ind = F.one_hot(near.long(), num_classes=2)
    d.protein_atom_feature = torch.cat(
        [F.one_hot(elem_idx, 6), F.one_hot(aa, 20), backbone, protein_arm_ind], dim=-1)
    d.protein_decomp_group_idx = torch.full((n_protein,), -1, dtype=torch.long)

    # FeaturizeLigandBond('fc') + uniform initial bond types
    d.ligand_fc_bond_index = fc_bond_index(n_lig)
    d.ligand_fc_bond_type = torch.randint(0, 5, (d.ligand_fc_bond_index.size(1),), generator=gen)

    # x_T and v_T (driver: sample_diffusion_decomp.py:163-176, 305-312)
    d.init_ligand_pos = centers[d.ligand_decomp_mask] + \
        torch.randn(n_lig, 3, generator=gen) * stds[d.ligand_decomp_mask]
    d.init_ligand_v = torch.randint(0, 8, (n_lig,), generator=gen)
    return d


def make_batch(n_pockets: int, n_protein=370, arm_sizes: Sequence[int] = (8, 8), n_scaffold: int = 14,
               seed: int = 0, dense: bool = False, n_full_extra: int = 0,
               ragged: bool = False) -> Dict[str, torch.Tensor]:
    """Collate `n_pockets` synthetic complexes and return the keyword arguments of
    `DecompScorePosNet3D.sample_diffusion` (CPU tensors).  `ragged=True` varies the sizes per pocket.
    `n_full_extra > 0` also builds the full-protein cloud used by the clash drift (cfg 3)."""
    gen = torch.Generator().manual_seed(seed)
    pockets = []
    for p in range(n_pockets):
        if ragged:
            npi = max(8, int(n_protein * (0.5 + torch.rand(1, generator=gen).item())))
            arms = [max(1, int(s + torch.randint(-3, 4, (1,), generator=gen).item())) for s in arm_sizes]
            nsc = max(0, int(n_scaffold + torch.randint(-5, 6, (1,), generator=gen).item()))
            if p % 5 == 4:
                nsc = 0  # a pocket without scaffold atoms
        else:
            npi = n_protein[p] if isinstance(n_protein, (list, tuple)) else n_protein
            arms, nsc = arm_sizes, n_scaffold
        pockets.append(make_pocket(gen, npi, arms, nsc, dense))
    return collate_pockets(pockets, n_full_extra, gen)


def collate_pockets(pockets: Sequence[ProteinLigandData], n_full_extra: int = 0,
                    gen: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """Collate complexes (`make_pocket`) into the keyword arguments of `DecompScorePosNet3D.sample_diffusion`."""
    n_pockets = len(pockets)
    batch = Batch.from_data_list(pockets, follow_batch=FOLLOW_BATCH)
    n_lig = [p.ligand_atom_mask.numel() for p in pockets]
    batch_ligand = torch.repeat_interleave(torch.arange(n_pockets), torch.tensor(n_lig))
    kw = dict(
        protein_pos=batch.protein_pos,
        protein_v=batch.protein_atom_feature.float(),
        batch_protein=batch.protein_element_batch,
        protein_group_idx=batch.protein_decomp_group_idx,
        init_ligand_pos=batch.init_ligand_pos,
        init_ligand_v=batch.init_ligand_v,
        ligand_v_aux=batch.ligand_atom_aux_feature.float(),
        batch_ligand=batch_ligand,
        ligand_group_idx=batch.ligand_decomp_group_idx,
        ligand_atom_mask=None,
        prior_centers=batch.ligand_decomp_centers,
        prior_stds=batch.ligand_decomp_stds,
        prior_num_atoms=batch.ligand_decomp_num_atoms,
        batch_prior=batch.ligand_decomp_centers_batch,
        prior_group_idx=batch.prior_group_idx,
        ligand_fc_bond_index=batch.ligand_fc_bond_index,
        init_ligand_fc_bond_type=batch.ligand_fc_bond_type,
        batch_ligand_bond=batch.ligand_fc_bond_type_batch,
        ligand_decomp_batch=batch.ligand_decomp_mask,
        ligand_decomp_index=batch.ligand_atom_mask,
    )
    if n_full_extra > 0:
        full_pos, full_batch = [], []
        for g, p in enumerate(pockets):
            extra = torch.randn(n_full_extra, 3, generator=gen) * 20.0
            fp = torch.cat([p.protein_pos, extra], 0)
            full_pos.append(fp)
            full_batch.append(torch.full((fp.size(0),), g, dtype=torch.long))
        kw['full_protein_pos'] = torch.cat(full_pos, 0)
        kw['full_batch_protein'] = torch.cat(full_batch, 0)
    return kw


def forward_kwargs(kw: Dict[str, torch.Tensor], time_step: Optional[torch.Tensor] = None) -> Dict:
    """Map `sample_diffusion` kwargs onto `forward` kwargs (decompdiff.py:578-599)."""
    return dict(
        protein_pos=kw['protein_pos'], protein_v=kw['protein_v'], batch_protein=kw['batch_protein'],
        protein_group_idx=kw['protein_group_idx'],
        init_ligand_pos=kw['init_ligand_pos'], init_ligand_v=kw['init_ligand_v'],
        init_ligand_v_aux=kw['ligand_v_aux'], batch_ligand=kw['batch_ligand'],
        ligand_group_idx=kw['ligand_group_idx'],
        prior_centers=kw['prior_centers'], prior_stds=kw['prior_stds'], batch_prior=kw['batch_prior'],
        prior_group_idx=kw['prior_group_idx'],
        ligand_fc_bond_index=kw['ligand_fc_bond_index'],
        init_ligand_fc_bond_type=kw['init_ligand_fc_bond_type'],
        ligand_atom_mask=kw.get('ligand_atom_mask'), time_step=time_step)


DEFAULT_MODEL_CONFIG = dict(  # `model:` of /root/reference/configs/training.yml:16-57
    model_mean_type='C0', beta_schedule='sigmoid', beta_start=1.e-7, beta_end=2.e-3,
    v_beta_schedule='cosine', v_beta_s=0.01, num_diffusion_timesteps=1000,
    v_mode='categorical', v_net_type='mlp', loss_pos_type='mse', sample_time_method='symmetric',
    bond_diffusion=True, bond_net_type='lin', num_bond_classes=5, prior_types=False,
    h_node_in_bond_net=True, add_prior_node=False, time_emb_dim=0, time_emb_mode='simple',
    center_pos_mode='protein', node_indicator=True, model_type='uni_o2_bond', num_blocks=1,
    num_layers=6, hidden_dim=128, n_heads=16, edge_feat_dim=4, num_r_gaussian=20, knn=32,
    act_fn='relu', norm=True, cutoff_mode='knn', r_max=10., x2h_out_fc=False, sync_twoup=False,
)
PROTEIN_FEATURE_DIM = 29   # 27 (FeaturizeProteinAtom) + 2 (AddDecompIndicator)
LIGAND_FEATURE_DIM = 10    # 8 ('basic' atom types) + 2 (arm/scaffold indicator)
NUM_CLASSES = 8


def synthetic_state_dict(model, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic random weights keyed by parameter name (no checkpoint offline).

    Linear weights ~ U(+-1/sqrt(fan_in)) as torch's default init, Linear biases ~ 0.05 N(0,1),
    LayerNorm affine = (1 + 0.1 N(0,1), 0.1 N(0,1)) so that every term of the arithmetic is exercised.
    Schedule tables and buffers keep their values.  Independent of module construction order, so the
    reference model, the oracle and the CUDA path can all be given the very same tensors.
    """
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(seed)
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    for name in sorted(sd):
        if name not in trainable:
            continue
        t = sd[name]
        if '.net.1.' in name:                      # LayerNorm
            r = 0.1 * torch.randn(t.shape, generator=gen)
            sd[name] = (1.0 + r) if name.endswith('weight') else r
        elif t.dim() >= 2:
            bound = 1.0 / (t.size(1) ** 0.5)
            sd[name] = (torch.rand(t.shape, generator=gen) * 2 - 1) * bound
        else:
            sd[name] = 0.05 * torch.randn(t.shape, generator=gen)
    return sd


def step_noise(n_ligand: int, n_bonds: int, num_steps: int, seed: int, num_classes: int = NUM_CLASSES,
               num_bond_classes: int = 5):
    """The draws a CPU run of the reference makes after `torch.manual_seed(seed)`: per step
    rand(n,C) -> rand(Eb,Cb) -> randn(n,3) (transitions.py:79 via decompdiff.py:620,633; :680)."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num_steps):
        out.append({'u_atom': torch.rand(n_ligand, num_classes, generator=gen),
                    'u_bond': torch.rand(n_bonds, num_bond_classes, generator=gen),
                    'eps_pos': torch.randn(n_ligand, 3, generator=gen)})
    return out


def make_raw_pocket(seed: int = 0, n_protein: int = 60, arm_sizes: Sequence[int] = (3, 4), n_scaffold: int = 5,
                    arm_radius: float = 10.0) -> ProteinLigandData:
    """A synthetic complex BEFORE the sampling-time transforms: the attributes a dataset item carries when the driver
    receives it (scripts/sample_diffusion_decomp.py:556-590) - protein atoms with element / residue / backbone
    fields, a reference ligand with its arm decomposition (`ligand_atom_mask`, -1 = scaffold) and the arm sub-pocket
    masks.  Feed it to `transforms.FeaturizeProteinAtom`, `prior.compute_golden_prior_from_data` (ref_prior) or
    `prior.substitute_golden_prior_with_given_prior` (beta_prior) and then to `sampling.sample_diffusion_ligand_decomp`."""
    gen = torch.Generator().manual_seed(seed)
    num_arms = len(arm_sizes)
    d = ProteinLigandData()
    d.protein_pos = torch.randn(n_protein, 3, generator=gen) * 8.0
    d.protein_element = PROTEIN_ELEMENTS[torch.randint(0, 6, (n_protein,), generator=gen)]
    d.protein_atom_to_aa_type = torch.randint(0, 20, (n_protein,), generator=gen)
    d.protein_is_backbone = torch.randint(0, 2, (n_protein,), generator=gen).bool()
    centers = torch.randn(num_arms + 1, 3, generator=gen) * 3.0
    spread = 0.5 + torch.rand(num_arms + 1, 1, generator=gen)
    mask: List[int] = []
    for a, s in enumerate(arm_sizes):
        mask += [a] * s
    mask += [-1] * n_scaffold
    d.ligand_atom_mask = torch.tensor(mask, dtype=torch.long)
    part = torch.where(d.ligand_atom_mask < 0, torch.tensor(num_arms), d.ligand_atom_mask)
    d.ligand_pos = centers[part] + torch.randn(len(mask), 3, generator=gen) * spread[part]
    d.ligand_element = torch.full((len(mask),), 6, dtype=torch.long)
    d.num_arms = num_arms
    d.num_scaffold = 1 if n_scaffold > 0 else 0
    d.pocket_atom_masks = torch.stack([(d.protein_pos - centers[a]).norm(dim=-1) < arm_radius for a in range(num_arms)])
    return d


def beta_prior_dict(seed: int, arm_sizes: Sequence[int] = (3, 4), n_scaffold: int = 5, scalar_scaffold_cov: bool = True) -> Dict:
    """A synthetic 'beta prior' in the pickle format of utils/prior.py:48-68: tuples (num, iso_mu, iso_cov, aniso_mu, aniso_cov);
    the scaffold variance may be a scalar (utils/transforms.py:229-236)."""
    rng = np.random.RandomState(seed)
    arms = [(int(n), rng.randn(3) * 3.0, np.eye(3) * float(rng.uniform(0.2, 2.0)), None, None) for n in arm_sizes]
    sca = []
    if n_scaffold > 0:
        var = float(rng.uniform(0.2, 2.0))
        sca.append((int(n_scaffold), rng.randn(3) * 3.0, var if scalar_scaffold_cov else np.eye(3) * var, None, None))
    return {'arms_prior': arms, 'scaffold_prior': sca, 'num_arms': len(arms), 'num_scaffold': len(sca)}


def select_pockets(kw: Dict[str, torch.Tensor], ids: Sequence[int]):
    """(sub_kw, rows): the `sample_diffusion` keyword arguments of a sub-batch holding pockets `ids` (ascending) of `kw`, re-numbered
    0..len(ids)-1 with the collate offsets of utils/data.py:439-444 re-applied.  Complexes never interact, so running a
    sub-batch gives the rows of the full batch that belong to those pockets (used for sharding and by the parity tests,
    which run the CPU oracle on a few pockets of a 64-pocket batch).  `rows['ligand']` / `rows['bond']` are the rows of the
    full batch's per-atom / per-bond tensors that the sub-batch holds."""
    ids = [int(i) for i in ids]
    if sorted(ids) != ids or len(set(ids)) != len(ids):
        raise ValueError('ids must be ascending and unique')
    B = int(kw['batch_protein'].max()) + 1
    new_id = torch.full((B,), -1, dtype=torch.long)
    new_id[torch.tensor(ids)] = torch.arange(len(ids))
    bp, bl, bb, bpr = kw['batch_protein'], kw['batch_ligand'], kw['batch_ligand_bond'], kw['batch_prior']
    mp, ml, mb, mpr = new_id[bp] >= 0, new_id[bl] >= 0, new_id[bb] >= 0, new_id[bpr] >= 0
    # old -> new row numbers of ligand atoms and prior rows (for the index-valued tensors)
    lig_new = torch.cumsum(ml.long(), 0) - 1
    prior_new = torch.cumsum(mpr.long(), 0) - 1
    out = dict(kw)
    out.update(
        protein_pos=kw['protein_pos'][mp], protein_v=kw['protein_v'][mp], batch_protein=new_id[bp[mp]],
        protein_group_idx=kw['protein_group_idx'][mp],
        init_ligand_pos=kw['init_ligand_pos'][ml], init_ligand_v=kw['init_ligand_v'][ml], ligand_v_aux=kw['ligand_v_aux'][ml],
        batch_ligand=new_id[bl[ml]], ligand_group_idx=kw['ligand_group_idx'][ml],
        ligand_atom_mask=None if kw.get('ligand_atom_mask') is None else kw['ligand_atom_mask'][ml],
        prior_centers=kw['prior_centers'][mpr], prior_stds=kw['prior_stds'][mpr], prior_num_atoms=kw['prior_num_atoms'][mpr],
        batch_prior=new_id[bpr[mpr]], prior_group_idx=kw['prior_group_idx'][mpr],
        ligand_fc_bond_index=lig_new[kw['ligand_fc_bond_index'][:, mb]],
        init_ligand_fc_bond_type=kw['init_ligand_fc_bond_type'][mb], batch_ligand_bond=new_id[bb[mb]],
        ligand_decomp_batch=prior_new[kw['ligand_decomp_batch'][ml]], ligand_decomp_index=kw['ligand_decomp_index'][ml],
    )
    if 'full_protein_pos' in kw:
        mf = new_id[kw['full_batch_protein']] >= 0
        out['full_protein_pos'] = kw['full_protein_pos'][mf]
        out['full_batch_protein'] = new_id[kw['full_batch_protein'][mf]]
    rows = {'ligand': ml.nonzero().squeeze(1), 'bond': mb.nonzero().squeeze(1)}
    return out, rows
