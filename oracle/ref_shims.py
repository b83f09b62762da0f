"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

Pure-torch stand-ins for the third-party wheels the reference hot path imports
but this image does not have (`torch_scatter`, `torch_sparse`, `torch_cluster`
via `torch_geometric.nn.knn_graph`, `torch_geometric.data`, `rdkit`, `openbabel`,
`easydict`).  With these installed into `sys.modules`, the reference's own
`models/decompdiff.py` imports and runs UNMODIFIED from `/root/reference`
(only in the build container; the GPU box has no `/root/reference`).

The unmodified reference run through these shims is what pins
`oracle/restate.py` (see `oracle/make_golden.py`) - the reference has no tests
or golden vectors of its own.

Semantics restated from the published behaviour of the libraries (not in tree):
* torch_scatter 2.1: scatter_sum / scatter_mean / scatter_softmax / scatter_min / scatter_max
* torch_cluster 1.6 `knn_graph(x, k, batch, loop=False, flow='source_to_target')`
* torch_sparse 0.6 `SparseTensor(row, col, value, sparse_sizes)` row-select / storage
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

import torch

REF_ROOT_CANDIDATES = ('/root/reference',)


# ----------------------------------------------------------------------------
# torch_scatter
# ----------------------------------------------------------------------------
def _bcast_index(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def _out_size(src, index, dim, dim_size):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = dim_size
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _bcast_index(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    total = scatter_sum(src, index, dim, out, dim_size)
    dim_size = total.size(dim)
    idx1 = index if index.dim() == 1 else index
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, idx1, 0 if index.dim() == 1 else dim, None, dim_size)
    count = count.clamp_(min=1)
    count = _bcast_index(count, total, dim if dim >= 0 else total.dim() + dim)
    if total.is_floating_point():
        return total.div_(count)
    return total.div_(count, rounding_mode='floor')


def _scatter_reduce(src, index, dim, dim_size, reduce):
    index_b = _bcast_index(index, src, dim)
    size = _out_size(src, index_b, dim, dim_size)
    fill = float('inf') if reduce == 'amin' else float('-inf')
    out = torch.full(size, fill, dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(dim, index_b, src, reduce=reduce, include_self=True)
    return out, index_b


def _arg_of(src, index_b, out, dim):
    # argument index of the reduced element (first match), torch_scatter returns src.size(dim) when empty
    n = src.size(dim)
    pos = torch.arange(n, device=src.device)
    shape = [1] * src.dim()
    shape[dim] = n
    pos = pos.view(shape).expand_as(src)
    hit = src == out.gather(dim, index_b)
    cand = torch.where(hit, pos, torch.full_like(pos, n))
    arg = torch.full(out.size(), n, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(dim, index_b, cand, reduce='amin', include_self=True)
    return arg


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    if dim < 0:
        dim = src.dim() + dim
    res, index_b = _scatter_reduce(src, index, dim, dim_size, 'amin')
    arg = _arg_of(src, index_b, res, dim)
    res = torch.where(torch.isinf(res) & (res > 0), torch.zeros_like(res), res)
    return res, arg


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    if dim < 0:
        dim = src.dim() + dim
    res, index_b = _scatter_reduce(src, index, dim, dim_size, 'amax')
    arg = _arg_of(src, index_b, res, dim)
    res = torch.where(torch.isinf(res) & (res < 0), torch.zeros_like(res), res)
    return res, arg


def scatter_softmax(src, index, dim=-1, eps=1e-12, dim_size=None):
    if dim < 0:
        dim = src.dim() + dim
    index_b = _bcast_index(index, src, dim)
    mx, _ = _scatter_reduce(src, index, dim, dim_size, 'amax')
    centred = src - mx.gather(dim, index_b)
    ex = centred.exp()
    total = scatter_sum(ex, index, dim, None, mx.size(dim))
    return ex / total.gather(dim, index_b)


# ----------------------------------------------------------------------------
# torch_geometric.nn.knn_graph (-> torch_cluster.knn)
# ----------------------------------------------------------------------------
def sq_dist_direct(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Pairwise squared distance by direct differences, fixed summation order
    ((dx*dx)+(dy*dy))+(dz*dz) in fp32 - the arithmetic the CUDA graph kernel reproduces
    bit-for-bit (no FMA contraction)."""
    dx = a[:, None, 0] - b[None, :, 0]
    dy = a[:, None, 1] - b[None, :, 1]
    dz = a[:, None, 2] - b[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def knn_graph(x, k, batch=None, loop=False, flow='source_to_target', cosine=False, num_workers=1):
    assert flow == 'source_to_target' and not cosine
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=x.device)
    srcs, dsts = [], []
    num_graphs = int(batch.max()) + 1 if n > 0 else 0
    counts = torch.bincount(batch, minlength=num_graphs).tolist()
    start = 0
    for g in range(num_graphs):
        m = counts[g]
        if m == 0:
            continue
        idx = torch.arange(start, start + m, device=x.device)
        assert bool((batch[idx] == g).all()), 'batch must be sorted'
        d2 = sq_dist_direct(x[idx], x[idx])
        if not loop:
            d2.fill_diagonal_(float('inf'))
        kk = min(k, m if loop else m - 1)
        if kk > 0:
            order = torch.sort(d2, dim=1, stable=True).indices[:, :kk]  # ties -> lower index
            srcs.append(idx[order].reshape(-1))
            dsts.append(idx[:, None].expand(m, kk).reshape(-1))
        start += m
    if not srcs:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(srcs), torch.cat(dsts)], dim=0)


def _unsupported(*a, **k):
    raise NotImplementedError('not on the hot path (reference radius mode is itself broken: '
                              'uni_transformer_edge.py:351 uses undefined self.r)')


# ----------------------------------------------------------------------------
# torch_sparse.SparseTensor (row-select + storage views only)
# ----------------------------------------------------------------------------
class _Storage:
    def __init__(self, row, col, value):
        self._row, self._col, self._value = row, col, value

    def row(self):
        return self._row

    def col(self):
        return self._col

    def value(self):
        return self._value


class SparseTensor:
    def __init__(self, row=None, col=None, value=None, sparse_sizes=None, _sorted=False):
        if not _sorted:
            key = row * int(sparse_sizes[1]) + col
            perm = torch.sort(key, stable=True).indices
            row, col = row[perm], col[perm]
            value = value[perm] if value is not None else None
        self._sizes = tuple(int(s) for s in sparse_sizes)
        self.storage = _Storage(row, col, value)
        counts = torch.bincount(row, minlength=self._sizes[0])
        self._rowptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])

    def __getitem__(self, index):
        assert torch.is_tensor(index) and index.dtype == torch.long and index.dim() == 1
        start = self._rowptr[index]
        length = self._rowptr[index + 1] - start
        total = int(length.sum())
        new_row = torch.repeat_interleave(torch.arange(index.numel(), device=index.device), length)
        offs = torch.arange(total, device=index.device) - torch.repeat_interleave(
            torch.cat([length.new_zeros(1), length.cumsum(0)[:-1]]), length)
        gather = torch.repeat_interleave(start, length) + offs
        col = self.storage._col[gather]
        val = self.storage._value[gather] if self.storage._value is not None else None
        return SparseTensor(row=new_row, col=col, value=val,
                            sparse_sizes=(index.numel(), self._sizes[1]), _sorted=True)

    def set_value(self, value, layout=None):
        return SparseTensor(row=self.storage._row, col=self.storage._col, value=value,
                            sparse_sizes=self._sizes, _sorted=True)

    def sum(self, dim):
        assert dim == 1
        row = self.storage._row
        if self.storage._value is None:
            src = torch.ones(row.numel(), device=row.device)
        else:
            src = self.storage._value
        return torch.zeros(self._sizes[0], dtype=src.dtype, device=row.device).scatter_add_(0, row, src)


# ----------------------------------------------------------------------------
# misc stubs
# ----------------------------------------------------------------------------
class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {}, **kwargs)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) else x for x in v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


class _Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_INSTALLED = False


def install():
    """Register the stand-in modules.  Idempotent."""
    global _INSTALLED
    if _INSTALLED:
        return
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from decompdiff_b200 import batch as _b  # PyG-free Data/Batch (the product's own stand-in)

    _module('torch_scatter', scatter_sum=scatter_sum, scatter_add=scatter_add, scatter_mean=scatter_mean,
            scatter_min=scatter_min, scatter_max=scatter_max, scatter_softmax=scatter_softmax)
    _module('torch_sparse', SparseTensor=SparseTensor)
    tg = _module('torch_geometric')
    tg.nn = _module('torch_geometric.nn', knn_graph=knn_graph, radius_graph=_unsupported,
                    radius=_unsupported, knn=_unsupported)
    tg.data = _module('torch_geometric.data', Data=_b.Data, Batch=_b.Batch)
    tg.loader = _module('torch_geometric.loader', DataLoader=type('DataLoader', (), {}))
    tg.transforms = _module('torch_geometric.transforms', Compose=_Compose)
    _module('easydict', EasyDict=EasyDict)
    for name in ('rdkit', 'rdkit.Chem', 'rdkit.Chem.rdchem', 'rdkit.Chem.AllChem', 'rdkit.Chem.Lipinski',
                 'rdkit.Chem.rdMolAlign', 'rdkit.Chem.ChemicalFeatures', 'rdkit.Chem.Descriptors',
                 'rdkit.Chem.rdMolTransforms', 'rdkit.Chem.QED', 'rdkit.Chem.Draw',
                 'rdkit.Geometry', 'rdkit.RDConfig', 'rdkit.RDLogger', 'rdkit.DataStructs',
                 'openbabel', 'openbabel.openbabel', 'openbabel.pybel', 'lmdb', 'sklearn.metrics.pairwise_'):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    _INSTALLED = True


def reference_root():
    for c in REF_ROOT_CANDIDATES:
        if os.path.isdir(os.path.join(c, 'models')):
            return c
    return None


def load_reference():
    """Import the unmodified reference packages (`models`, `utils`).  Returns the root or None."""
    root = reference_root()
    if root is None:
        return None
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    import models.decompdiff  # noqa: F401  (reference module, unmodified)
    return root


def reference_model_config():
    """`model:` section of /root/reference/configs/training.yml (the shipped configuration)."""
    import yaml
    root = reference_root()
    with open(os.path.join(root, 'configs', 'training.yml')) as f:
        return EasyDict(yaml.safe_load(f)['model'])
