"""Decomposed priors: per-arm / scaffold Gaussian statistics.

Host-side set-up that runs once per pocket before sampling; mirrors the functions of
/root/reference/utils/prior.py that the sampling driver calls
(scripts/sample_diffusion_decomp.py:568-578):

    compute_golden_prior_from_data            utils/prior.py:126-159   ('ref_prior')
    substitute_golden_prior_with_given_prior  utils/prior.py:70-88
    substitute_golden_prior_with_beta_prior   utils/prior.py:48-68     ('beta_prior', from a pickle)
    apply_std_coef / apply_num_atoms_change   utils/prior.py:91-123

A prior entry is the reference's 5-tuple `(num_atoms, iso_mu, iso_cov, aniso_mu, aniso_cov)`.
`NumAtomsSampler` (sklearn regressors loaded from a pickle, `num_atoms_mode='stat'`) is out of scope.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch
import torch.nn.functional as F

POCKET_CONTACT_THRESHOLD = 6.0      # utils/prior.py:129


def _matmul_like(a, b):
    """np.matmul as the reference calls it; torch inputs come back as torch tensors (what `__array_wrap__` used to do)."""
    if torch.is_tensor(a):
        return torch.from_numpy(np.matmul(a.numpy(), b.numpy()))
    return np.matmul(a, b)


def _eye_like(pos):
    return torch.from_numpy(np.eye(3)) if torch.is_tensor(pos) else np.eye(3)      # float64, as np.eye in the reference


def isotropic_covariance(pos):
    """sigma^2 I with sigma^2 = sum |x - mu|^2 / (3 n)  (utils/prior.py:11-20)."""
    assert len(pos.shape) == 2 and pos.shape[1] == 3
    centred = (pos - pos.mean(axis=0, keepdims=True)).reshape(-1, 1)
    return _matmul_like(centred.T, centred) / centred.shape[0] * _eye_like(pos)


def anisotropic_covariance(pos):
    """(x - mu)^T (x - mu) / n  (utils/prior.py:23-31)."""
    assert len(pos.shape) == 2 and pos.shape[1] == 3
    centred = pos - pos.mean(axis=0, keepdims=True)
    return _matmul_like(centred.T, centred) / centred.shape[0]


def get_iso_aniso_mu_cov(pos):
    """(iso_mu, iso_cov, aniso_mu, aniso_cov) of a set of atoms; empty sets give empty statistics (utils/prior.py:34-45)."""
    if pos.shape[0] == 0:
        return np.zeros_like(pos), np.eye(0), np.zeros_like(pos), np.eye(0)
    mu = pos.mean(axis=0)
    return mu, isotropic_covariance(pos), mu, anisotropic_covariance(pos)


def compute_golden_prior_from_data(data):
    """'ref_prior': statistics of the reference ligand's own arms / scaffold + 6 A contact masks of the pocket."""
    masks, arms = [], []
    for arm in range(data.num_arms):
        atoms = data.ligand_pos[data.ligand_atom_mask == arm, :]
        iso_mu, iso_cov, aniso_mu, aniso_cov = get_iso_aniso_mu_cov(atoms)
        arms.append((atoms.shape[0], iso_mu, iso_cov, aniso_mu, aniso_cov))
        masks.append(F.pairwise_distance(iso_mu.unsqueeze(0), data.protein_pos) < POCKET_CONTACT_THRESHOLD)
    scaffold = []
    atoms = data.ligand_pos[data.ligand_atom_mask == -1, :]
    if atoms.shape[0] > 0:
        iso_mu, iso_cov, aniso_mu, aniso_cov = get_iso_aniso_mu_cov(atoms)
        scaffold.append((atoms.shape[0], iso_mu, iso_cov, aniso_mu, aniso_cov))
        masks.append((F.pairwise_distance(iso_mu.unsqueeze(0), data.protein_pos) < POCKET_CONTACT_THRESHOLD).bool())
    data.scaffold_prior, data.arms_prior = scaffold, arms
    assert len(arms) == data.num_arms and len(scaffold) == data.num_scaffold
    data.pocket_prior_masks = torch.stack(masks)
    assert len(data.pocket_prior_masks) == data.num_arms + data.num_scaffold
    return data


def substitute_golden_prior_with_given_prior(data, prior_dict, protein_ligand_dist_th: float = 10.0):
    """Install `{arms_prior: [...], scaffold_prior: [...]}` on `data`; arm pockets = protein atoms within the threshold."""
    assert len(prior_dict['scaffold_prior']) <= 1
    data.num_arms, data.num_scaffold = len(prior_dict['arms_prior']), len(prior_dict['scaffold_prior'])
    data.arms_prior, data.scaffold_prior, pocket_masks = [], [], []
    for count, mu, cov, _, _ in prior_dict['arms_prior']:
        mu_t = torch.tensor(mu).float()
        data.arms_prior.append((count, mu_t, torch.tensor(cov).float(), None, None))
        # sklearn's pairwise_distances (euclidean) in the reference; direct differences in float64 agree to the last bit
        # on everything but exact-threshold ties
        dist = np.sqrt(((np.asarray(data.protein_pos, dtype=np.float64) - mu_t.reshape(1, 3).double().numpy()) ** 2).sum(-1))
        pocket_masks.append(dist < protein_ligand_dist_th)
    for count, mu, cov, _, _ in prior_dict['scaffold_prior']:
        data.scaffold_prior.append((count, torch.tensor(mu).float(), torch.tensor(cov).float(), None, None))
    n_protein = len(data.protein_pos)
    data.pocket_atom_masks = (torch.tensor(np.array(pocket_masks)).reshape(len(pocket_masks), n_protein)
                              if pocket_masks else torch.zeros(0, n_protein, dtype=torch.bool))
    return data


def substitute_golden_prior_with_beta_prior(data, beta_prior_path, protein_ligand_dist_th: float = 10.0):
    """'beta_prior': the same from a pickle `{arms_prior, scaffold_prior, num_arms, num_scaffold}` (utils/prior.py:48-68)."""
    with open(beta_prior_path, 'rb') as f:
        beta_prior = pickle.load(f)
    assert len(beta_prior['arms_prior']) == beta_prior['num_arms']
    assert len(beta_prior['scaffold_prior']) == beta_prior['num_scaffold']
    return substitute_golden_prior_with_given_prior(data, beta_prior, protein_ligand_dist_th)


def _rescaled(entry, std_coef=None, num_atoms_change=None):
    count, mu, cov = entry[:3]
    if std_coef is not None:
        cov *= std_coef ** 2               # in place, as the reference (utils/prior.py:95)
    if num_atoms_change is not None:
        count = max(count + num_atoms_change, 1)
    return (count, mu, cov, None, None)


def apply_std_coef(data, std_coef):
    data.arms_prior = [_rescaled(e, std_coef=std_coef) for e in data.arms_prior]
    assert len(data.scaffold_prior) <= 1
    data.scaffold_prior = [_rescaled(e, std_coef=std_coef) for e in data.scaffold_prior]


def apply_num_atoms_change(data, num_atoms_change):
    data.arms_prior = [_rescaled(e, num_atoms_change=num_atoms_change) for e in data.arms_prior]
    assert len(data.scaffold_prior) <= 1
    data.scaffold_prior = [_rescaled(e, num_atoms_change=num_atoms_change) for e in data.scaffold_prior]


class NumAtomsSampler:
    """Atom counts / prior stds predicted from pocket statistics (utils/prior.py:162-208, `num_atoms_mode='stat'`).
    `pred_models_dict` = {'arm_model', 'armstd_model', 'sca_model', 'scastd_model'}: any regressors with a scikit-learn style
    `.predict(X)` (the reference unpickles sklearn models).  Draws use `np.random` in the reference's order."""

    RADII = np.linspace(1, 10, 50)

    def __init__(self, pred_models_dict):
        self.arm_model = pred_models_dict['arm_model']
        self.armstd_model = pred_models_dict['armstd_model']
        self.sca_model = pred_models_dict['sca_model']
        self.scastd_model = pred_models_dict['scastd_model']

    @classmethod
    def _shell_counts(cls, centers, protein_pos):
        """(n_centers, 50) number of protein atoms within r of each centre, r = 1..10 A (:171-172)."""
        d = torch.norm(centers.view(-1, 1, 3) - protein_pos.view(1, -1, 3), p=2, dim=-1)
        return torch.stack([(d < r).sum(1) for r in cls.RADII], dim=1).numpy()

    def sample_arm_natoms(self, arm_centers, protein_pos):
        y = self.arm_model.predict(self._shell_counts(arm_centers, protein_pos))
        arm_natoms = self.sample_natoms_from_prediction(y, std=0.2)
        arm_stds = self.armstd_model.predict(arm_natoms[:, None])
        arm_stds = torch.from_numpy(np.asarray(arm_stds).astype(np.float32)).reshape(-1, 1).expand(-1, 3)
        return arm_natoms.tolist(), arm_stds

    def sample_sca_natoms(self, sca_center, arm_centers, arm_stds, protein_pos):
        feat = self._shell_counts(sca_center, protein_pos)
        dist = torch.norm(sca_center.view(-1, 1, 3) - arm_centers.view(1, -1, 3), p=2, dim=-1).numpy()
        res = [d - r for d, r in zip(dist, arm_stds.numpy())]            # zip over the single scaffold row, as the reference (:188)
        x = np.concatenate([feat, np.array([d.sum() for d in res])[:, None]], axis=-1)
        y = self.sca_model.predict(x)
        sca_natoms = self.sample_natoms_from_prediction(y, std=0.)
        sca_stds = self.scastd_model.predict(sca_natoms[:, None])
        assert len(sca_natoms) == len(sca_stds) == 1
        return sca_natoms.tolist()[0], torch.from_numpy(np.asarray(sca_stds).astype(np.float32)).expand(3)

    @staticmethod
    def sample_natoms_from_prediction(n, std, min_natoms=2):
        natoms = np.ceil(n + std * n * np.random.randn(len(n))).astype(int)
        return np.maximum(natoms, min_natoms)
