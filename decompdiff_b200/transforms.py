"""Sampling-time transforms of the decomposed-prior driver.

Host-side, once per sample, never per step.  Each class mirrors the call contract of the
reference transform of the same name (/root/reference/utils/transforms.py) - same constructor
arguments, same attributes written onto the `ProteinLigandData` object, same dtypes - so the
reference's `init_transform = Compose([...])` list (scripts/sample_diffusion_decomp.py:519-534)
can be built from this module unchanged.

    FeaturizeProteinAtom        utils/transforms.py:114-131   27-dim protein atom features
    FeaturizeLigandAtom         utils/transforms.py:134-160   only the constants the driver reads
    ComputeLigandAtomNoiseDist  utils/transforms.py:166-254   mu_k / sigma_k / atom counts per arm + scaffold
    AddDecompIndicator          utils/transforms.py:257-320   decomp mask, aux features, protein arm bit
    FeaturizeLigandBond         utils/transforms.py:323-391   directed bond index of the ligand graph
    get_atomic_number_from_index / is_aromatic_from_index      utils/transforms.py:73-96
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

MIN_PRIOR_STD = 0.6          # utils/transforms.py:195

# ligand atom-type vocabularies (utils/transforms.py:41-68): index -> atomic number [, aromatic]
_BASIC_ATOMIC_NUMBERS = (1, 6, 7, 8, 9, 15, 16, 17)
_AROMATIC_TYPES = ((1, False), (6, False), (6, True), (7, False), (7, True), (8, False), (8, True),
                   (9, False), (15, False), (15, True), (16, False), (16, True), (17, False))
_FULL_TYPES = ((1, 'S', False), (6, 'SP', False), (6, 'SP2', False), (6, 'SP2', True), (6, 'SP3', False),
               (7, 'SP', False), (7, 'SP2', False), (7, 'SP2', True), (7, 'SP3', False), (8, 'SP2', False),
               (8, 'SP2', True), (8, 'SP3', False), (9, 'SP3', False), (15, 'SP2', False), (15, 'SP2', True),
               (15, 'SP3', False), (15, 'SP3D', False), (16, 'SP2', False), (16, 'SP2', True), (16, 'SP3', False),
               (16, 'SP3D', False), (16, 'SP3D2', False), (17, 'SP3', False))
NUM_ATOM_TYPES = {'basic': len(_BASIC_ATOMIC_NUMBERS), 'add_aromatic': len(_AROMATIC_TYPES), 'full': len(_FULL_TYPES)}


def _as_list(index) -> List[int]:
    return [int(i) for i in (index.tolist() if hasattr(index, 'tolist') else index)]


def get_atomic_number_from_index(index, mode: str) -> List[int]:
    """Atom-type index -> atomic number (utils/transforms.py:73-82); unknown mode raises ValueError."""
    if mode == 'basic':
        return [_BASIC_ATOMIC_NUMBERS[i] for i in _as_list(index)]
    if mode == 'add_aromatic':
        return [_AROMATIC_TYPES[i][0] for i in _as_list(index)]
    if mode == 'full':
        return [_FULL_TYPES[i][0] for i in _as_list(index)]
    raise ValueError(mode)


def is_aromatic_from_index(index, mode: str):
    """Aromatic flag per atom, or None in 'basic' mode (utils/transforms.py:85-94)."""
    if mode == 'add_aromatic':
        return [_AROMATIC_TYPES[i][1] for i in _as_list(index)]
    if mode == 'full':
        return [_FULL_TYPES[i][2] for i in _as_list(index)]
    if mode == 'basic':
        return None
    raise ValueError(mode)


class Compose:
    """`torch_geometric.transforms.Compose`: apply the transforms in order."""

    def __init__(self, transforms: Sequence[Callable]):
        self.transforms = list(transforms)

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data


class FeaturizeProteinAtom:
    """[element one-hot over (H,C,N,O,S,Se) | residue one-hot (20) | backbone bit] -> `protein_atom_feature` (27, int64)."""

    def __init__(self):
        self.atomic_numbers = torch.LongTensor([1, 6, 7, 8, 16, 34])
        self.max_num_aa = 20

    @property
    def protein_feature_dim(self) -> int:
        return self.atomic_numbers.numel() + self.max_num_aa + 1

    def __call__(self, data):
        is_element = data.protein_element.view(-1, 1) == self.atomic_numbers.view(1, -1)
        residue = F.one_hot(data.protein_atom_to_aa_type, num_classes=self.max_num_aa)
        backbone = data.protein_is_backbone.view(-1, 1).long()
        data.protein_atom_feature = torch.cat([is_element, residue, backbone], dim=-1)
        return data


class FeaturizeLigandAtom:
    """Only what the sampling driver reads: `ligand_feature_dim` (= num_classes) and the optional type priors
    (utils/transforms.py:136-155).  Featurising reference ligands needs RDKit fields and is out of scope."""

    def __init__(self, mode: str = 'basic', prior_types: bool = True):
        if mode not in NUM_ATOM_TYPES:
            raise AssertionError(mode)
        self.mode, self.prior_types = mode, prior_types
        if mode == 'basic' and prior_types:
            self.atom_types_prob = np.array([0., 0.6716, 0.1174, 0.1689, 0.01315, 0.01117, 0.01128, 0.00647])
            self.bond_types_prob = np.array([0.9170, 0.0433, 0.00687, 0.000173, 0.03266])
        else:
            self.atom_types_prob = self.bond_types_prob = None

    @property
    def ligand_feature_dim(self) -> int:
        return NUM_ATOM_TYPES[self.mode]


def _iso_std(cov) -> torch.Tensor:
    """sqrt of the isotropic variance, accepting a (3,3) matrix or (beta priors) a scalar."""
    cov = torch.as_tensor(cov)
    var = cov if cov.dim() == 0 else cov[0, 0]
    return torch.sqrt(var)


def _row3(value) -> torch.Tensor:
    return torch.tensor([float(value)]).reshape(1, 1).expand(-1, 3).clone()


class ComputeLigandAtomNoiseDist:
    """Prior centres mu_k, standard deviations sigma_k (one row per arm, then the scaffold row) and atom counts."""

    def __init__(self, version: str):
        if version not in ('subpocket', 'ref_prior', 'beta_prior'):
            raise AssertionError(version)
        self.version = version

    def __call__(self, data):
        if self.version == 'subpocket':
            centers = []
            for arm, mask in enumerate(data.pocket_atom_masks):
                if mask.sum() > 0:
                    centers.append(data.protein_pos[mask].mean(0))
                else:                      # an arm with no protein atom around it falls back to its ligand atoms
                    centers.append(data.ligand_pos[data.ligand_atom_mask == arm].mean(0))
            data.arm_centers = torch.stack(centers)
            data.scaffold_center = data.protein_pos.mean(0).unsqueeze(0)
            data.ligand_decomp_centers = torch.cat([data.arm_centers, data.scaffold_center], dim=0)
            data.ligand_decomp_stds = torch.ones_like(data.ligand_decomp_centers)
        else:
            centers, stds = [], []
            for arm in range(data.num_arms):
                count, mu, cov = data.arms_prior[arm][:3]
                std = torch.clamp(_row3(_iso_std(cov)), min=MIN_PRIOR_STD) if count > 1 else _row3(MIN_PRIOR_STD)
                centers.append(torch.as_tensor(mu).unsqueeze(0))
                stds.append(std)
            if len(data.scaffold_prior) > 0:
                assert len(data.scaffold_prior) == 1 and len(data.scaffold_prior) == data.num_scaffold
                count, mu, cov = data.scaffold_prior[0][:3]
                centers.append(torch.as_tensor(mu).unsqueeze(0))
                if self.version == 'ref_prior':
                    if count > 1:
                        stds.append(torch.clamp(_row3(_iso_std(cov)), min=MIN_PRIOR_STD))
                    elif count == 1:
                        stds.append(_row3(MIN_PRIOR_STD))
                    else:
                        raise ValueError(f'scaffold_atom_num = {count}, not valid!')
                else:                       # beta_prior: the variance may be a scalar; clamped all the same
                    stds.append(torch.clamp(_row3(_iso_std(cov)), min=MIN_PRIOR_STD))
            else:                           # no scaffold: the row still exists (protein centroid, minimum std)
                centers.append(data.protein_pos.mean(0).unsqueeze(0))
                stds.append(_row3(MIN_PRIOR_STD))
            data.ligand_decomp_centers = torch.cat(centers, dim=0)
            data.ligand_decomp_stds = torch.cat(stds, dim=0).float()

        data.arm_num_atoms = torch.tensor([(data.ligand_atom_mask == a).sum() for a in range(data.num_arms)])
        data.scaffold_num_atoms = (data.ligand_atom_mask == -1).sum().unsqueeze(0)
        data.ligand_decomp_num_atoms = torch.cat([data.arm_num_atoms, data.scaffold_num_atoms])
        return data


class AddDecompIndicator:
    """`ligand_decomp_mask` (scaffold -1 -> num_arms), 2-dim arm/scaffold aux feature, protein arm-pocket bit."""

    def __init__(self, max_num_arms: int = 10, global_prior_index=None, add_ord_feat: bool = False,
                 add_to_protein: bool = True, add_to_ligand: bool = True):
        self.max_num_arms = max_num_arms
        self.global_prior_index = global_prior_index
        self.add_ord_feat = add_ord_feat
        self.num_classes = max_num_arms + 1
        self.add_to_protein, self.add_to_ligand = add_to_protein, add_to_ligand

    @property
    def protein_feature_dim(self) -> int:
        return 2 + (self.num_classes if self.add_ord_feat else 0)

    ligand_feature_dim = protein_feature_dim

    def __call__(self, data):
        data.prior_group_idx = torch.arange(data.num_arms + 1, dtype=torch.long)
        data.max_decomp_group = self.num_classes
        if self.add_to_ligand:
            decomp = data.ligand_atom_mask.clone()
            decomp[decomp == -1] = data.num_arms
            data.ligand_decomp_mask = decomp
            data.ligand_decomp_group_idx = decomp            # the reference aliases the same tensor (:290)
            in_arm = F.one_hot((data.ligand_atom_mask >= 0).long(), num_classes=2)
            if self.add_ord_feat:
                data.ligand_atom_aux_feature = torch.cat([F.one_hot(decomp, self.num_classes), in_arm], -1)
            else:
                data.ligand_atom_aux_feature = in_arm
        if self.add_to_protein:
            in_arm_pocket = F.one_hot((data.pocket_atom_masks.sum(0) > 0).long(), num_classes=2)
            # the reference's per-arm one-hot is written through a chained index (`x[mask][arm] = 1`, :309), i.e. into a
            # temporary: it stays all zero - reproduced as such
            ordinal = torch.zeros([len(data.protein_pos), self.num_classes])
            data.protein_decomp_group_idx = torch.full((len(data.protein_pos),), -1, dtype=torch.long)
            extra = [ordinal, in_arm_pocket] if self.add_ord_feat else [in_arm_pocket]
            data.protein_atom_feature = torch.cat([data.protein_atom_feature] + extra, -1)
        return data


def _pairs_without_self(dst_atoms: torch.Tensor, src_atoms: torch.Tensor):
    """All (src, dst) with dst-major order and src != dst."""
    dst = torch.repeat_interleave(dst_atoms, src_atoms.numel())
    src = src_atoms.repeat(dst_atoms.numel())
    keep = dst != src
    return src[keep], dst[keep]


class FeaturizeLigandBond:
    """Directed ligand bond graph: 'fc' all ordered pairs; 'decomp_fc' pairs inside each arm / the scaffold;
    'scaffold_fc' pairs inside each arm plus every scaffold atom to every atom."""

    def __init__(self, mode: str = 'fc', set_bond_type: bool = False):
        self.mode, self.set_bond_type = mode, set_bond_type

    def __call__(self, data):
        n_atoms = len(data.ligand_atom_mask)          # the driver resets only ligand_atom_mask when sampling
        everyone = torch.arange(n_atoms)
        if self.mode == 'fc':
            src, dst = _pairs_without_self(everyone, everyone)
        elif self.mode == 'decomp_fc':
            parts = []
            for g in range(data.num_arms + data.num_scaffold):
                members = (data.ligand_decomp_mask == g).nonzero()[:, 0]
                parts.append(_pairs_without_self(members, members))
            src, dst = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
        elif self.mode == 'scaffold_fc':
            parts = []
            for g in range(data.num_arms):
                members = (data.ligand_decomp_mask == g).nonzero()[:, 0]
                parts.append(_pairs_without_self(members, members))
            scaffold = (data.ligand_atom_mask == -1).nonzero()[:, 0]
            parts.append(_pairs_without_self(everyone, scaffold))
            src, dst = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
        else:
            raise ValueError(self.mode)
        data.ligand_fc_bond_index = torch.stack([src, dst], dim=0)
        if self.set_bond_type and hasattr(data, 'ligand_bond_index'):
            n = data.ligand_pos.size(0)
            table = torch.zeros(n, n).long()
            table[data.ligand_bond_index[0], data.ligand_bond_index[1]] = data.ligand_bond_type
            data.ligand_fc_bond_type = table[data.ligand_fc_bond_index[0], data.ligand_fc_bond_index[1]]
        return data
