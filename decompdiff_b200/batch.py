"""PyG-free stand-ins for the batch container the sampling driver reads.

The reference collates `ProteinLigandData` objects with
`torch_geometric.data.Batch.from_data_list(..., follow_batch=FOLLOW_BATCH)`
(/root/reference/scripts/sample_diffusion_decomp.py:314-316) and relies on the
custom `__inc__` offsets of /root/reference/utils/data.py:439-444.  PyG is not
part of this image, so this module restates the small part of that contract the
hot path needs:

* `Data`      - attribute bag with `clone()`, item access and `keys`.
* `ProteinLigandData.__inc__` - the per-key collate increments.
* `Batch.from_data_list`      - concatenate tensors (dim -1 for `*index*` keys,
  dim 0 otherwise), add increments, emit `<key>_batch` / `<key>_ptr` for the
  keys listed in `follow_batch`, keep non-tensor attributes as python lists.
"""
from __future__ import annotations

import copy
from typing import Iterable, List, Sequence

import torch

# /root/reference/datasets/pl_data.py:11
FOLLOW_BATCH = ('protein_element', 'ligand_element', 'ligand_decomp_centers', 'ligand_fc_bond_type')


class Data:
    """Minimal attribute container (the part of `torch_geometric.data.Data` used by the path)."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    # -- mapping protocol -------------------------------------------------
    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.__dict__

    @property
    def keys(self):
        return [k for k in self.__dict__.keys() if not k.startswith('_')]

    def to_dict(self):
        return {k: self.__dict__[k] for k in self.keys}

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
        return out

    def to(self, device, non_blocking: bool = False):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device, non_blocking=non_blocking)
        return self

    # -- collate hooks ------------------------------------------------------
    def __cat_dim__(self, key, value, *args, **kwargs):
        # PyG rule: keys matching (index|face) are concatenated along the last dim.
        return -1 if ('index' in key or 'face' in key) else 0

    def __inc__(self, key, value, *args, **kwargs):
        return 0

    def __repr__(self):
        parts = []
        for k in self.keys:
            v = self.__dict__[k]
            parts.append(f'{k}={list(v.shape)}' if torch.is_tensor(v) else f'{k}={type(v).__name__}')
        return f'{self.__class__.__name__}({", ".join(parts)})'


class ProteinLigandData(Data):
    """Collate increments of /root/reference/utils/data.py:371-446 (decomp keys only)."""

    def __inc__(self, key, value, *args, **kwargs):
        if key == 'ligand_bond_index':
            return self['ligand_element'].size(0)
        if key == 'ligand_decomp_mask':
            # centres / num_atoms always occupy num_arms + 1 rows (utils/data.py:440)
            return self['num_arms'] + 1
        if key in ('ligand_decomp_group_idx', 'protein_decomp_group_idx'):
            return self['max_decomp_group']
        if key in ('ligand_fc_bond_index', 'ligand_full_bond_index'):
            return self['ligand_atom_mask'].size(0)
        return 0


class Batch(Data):
    """`Batch.from_data_list` restated for flat tensor attributes."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data], follow_batch: Iterable[str] = (),
                       exclude_keys: Iterable[str] = ()):
        follow_batch = set(follow_batch or ())
        exclude = set(exclude_keys or ())
        if len(data_list) == 0:
            raise ValueError('empty data list')
        keys: List[str] = [k for k in data_list[0].keys if k not in exclude]
        out = cls()
        out.num_graphs = len(data_list)
        for key in keys:
            values = [d[key] for d in data_list]
            first = values[0]
            if torch.is_tensor(first) and first.dim() > 0:
                cat_dim = data_list[0].__cat_dim__(key, first)
                pieces, inc = [], 0
                sizes = []
                for d, v in zip(data_list, values):
                    step = d.__inc__(key, v)
                    if isinstance(step, torch.Tensor):
                        step = int(step.item())
                    pieces.append(v + inc if (inc != 0 and v.dtype != torch.bool) else v)
                    sizes.append(v.size(cat_dim))
                    inc += step
                out[key] = torch.cat(pieces, dim=cat_dim)
                if key in follow_batch:
                    counts = torch.tensor(sizes, dtype=torch.long)
                    out[f'{key}_batch'] = torch.repeat_interleave(torch.arange(len(values)), counts)
                    out[f'{key}_ptr'] = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
            elif torch.is_tensor(first):
                out[key] = torch.stack(values)
            elif isinstance(first, (int, float)) and not isinstance(first, bool):
                out[key] = torch.tensor(values)
            else:
                out[key] = values
        return out

    @classmethod
    def from_replicas(cls, proto: Data, n: int, follow_batch: Iterable[str] = (), exclude_keys: Iterable[str] = (),
                      overrides=None, device=None):
        """`from_data_list([proto_0 .. proto_{n-1}])` for n samples that share every attribute of `proto` except the keys in
        `overrides` ({key: list of n tensors}) - the mini-batches of the sampling driver when the atom counts are fixed (same
        ligand_atom_mask for every sample, so the transforms give the same result).  Built with a handful of tensor ops on
        `device` instead of a per-sample python loop: only ONE copy of the shared tensors crosses to the device.  The result is
        bit-identical to `from_data_list` (tests/test_host_logic.py)."""
        follow_batch, exclude, overrides = set(follow_batch or ()), set(exclude_keys or ()), dict(overrides or {})
        out = cls()
        out.num_graphs = n
        dev = device
        for key in [k for k in proto.keys if k not in exclude]:
            first = proto[key]
            if torch.is_tensor(first) and first.dim() > 0:
                cat_dim = proto.__cat_dim__(key, first)
                step = proto.__inc__(key, first)
                step = int(step.item()) if isinstance(step, torch.Tensor) else int(step)
                if key in overrides:
                    vals = [v.to(dev) if dev is not None else v for v in overrides[key]]
                    if step != 0:
                        vals = [v + step * k if v.dtype != torch.bool else v for k, v in enumerate(vals)]
                    value = torch.cat(vals, dim=cat_dim)
                    size = vals[0].size(cat_dim)
                else:
                    v = first.to(dev) if dev is not None else first
                    size = v.size(cat_dim)
                    reps = [1] * v.dim()
                    reps[cat_dim] = n
                    value = v.repeat(*reps)
                    if step != 0 and v.dtype != torch.bool:
                        offs = (torch.arange(n, device=value.device, dtype=value.dtype) * step).repeat_interleave(size)
                        shape = [1] * v.dim()
                        shape[cat_dim] = n * size
                        value = value + offs.view(*shape)
                out[key] = value
                if key in follow_batch:
                    counts = torch.full((n,), size, dtype=torch.long, device=value.device)
                    out[f'{key}_batch'] = torch.repeat_interleave(torch.arange(n, device=value.device), counts)
                    out[f'{key}_ptr'] = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
            elif torch.is_tensor(first):
                v = first.to(dev) if dev is not None else first
                out[key] = torch.stack([v] * n)
            elif isinstance(first, (int, float)) and not isinstance(first, bool):
                t = torch.tensor([first] * n)
                out[key] = t.to(dev) if dev is not None else t
            else:
                out[key] = [first] * n
        return out
