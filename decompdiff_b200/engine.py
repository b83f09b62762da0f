"""Thin object wrappers over the C ABI (include/decompdiff_b200.h).

`EngineModel`  <-> ddb_model : weights handed over under their reference state_dict names.
`EngineBatch`  <-> ddb_batch : static topology + workspace + evolving (x_t, v_t, b_t) state.

torch is used here for device memory and streams only; every computation is a kernel of the library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib


def _stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _on_device(fn):
    """Run a method with the object's CUDA device current (kernel launches go to the thread's current device)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapped


def _host(t: torch.Tensor, dtype) -> torch.Tensor:
    return t.detach().to(device='cpu', dtype=dtype).contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('decompdiff_b200 needs a CUDA device (sm_100a); there is no CPU fallback')


def _device_of(*tensors, default=None) -> torch.device:
    """The CUDA device the call runs on: the first CUDA tensor among the inputs (as the reference, which computes where its
    tensors live), else `default`, else the current device."""
    for t in tensors:
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    if default is not None:
        return torch.device(default)
    return torch.device('cuda', torch.cuda.current_device())


class EngineModel:
    def __init__(self, cfg: Dict[str, int], state_dict: Dict[str, torch.Tensor], device=None, refine_only: bool = False,
                 cutoff_mode: str = 'knn', r_max: float = 0.0, mean_type: str = 'C0', time_emb=None):
        require_cuda()
        L = _lib.lib()
        self.device = _device_of(default=device)
        self.refine_only = refine_only
        if cutoff_mode not in ('knn', 'radius', 'hybrid'):
            raise ValueError(f'Not supported cutoff mode: {cutoff_mode}')      # uni_transformer_edge.py:358
        self.cutoff = ({'knn': 0, 'radius': 1, 'hybrid': 2}[cutoff_mode], float(r_max))
        if mean_type not in ('C0', 'noise'):
            raise ValueError(mean_type)      # models/decompdiff.py:610
        self.mean_noise = mean_type == 'noise'
        if time_emb not in (None, 'simple'):
            raise NotImplementedError(time_emb)      # models/decompdiff.py:182
        self.time_emb = time_emb
        with torch.cuda.device(self.device):
            self._init(L, cfg, state_dict)

    def _init(self, L, cfg, state_dict):
        c = _lib.Config(**cfg)
        self._h = C.c_void_p()
        _lib.check(L.ddb_model_create(C.byref(self._h), C.byref(c)))
        self.cfg = dict(cfg)
        for name, t in state_dict.items():
            if not torch.is_floating_point(t):
                continue
            ht = _host(t, torch.float32)
            _lib.check(L.ddb_model_set_tensor(self._h, name.encode(), C.c_void_p(ht.data_ptr()), ht.numel()))
        if self.refine_only:
            _lib.check(L.ddb_model_set_refine_only(self._h, 1))
        _lib.check(L.ddb_model_set_cutoff(self._h, self.cutoff[0], self.cutoff[1]))
        _lib.check(L.ddb_model_set_mean_type(self._h, 1 if self.mean_noise else 0))
        _lib.check(L.ddb_model_set_time_emb(self._h, 1 if self.time_emb == 'simple' else 0))
        _lib.check(L.ddb_model_finalize(self._h))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            try:
                _lib.lib().ddb_model_destroy(h)
            except Exception:
                pass
            self._h = None


class EngineBatch:
    """One collated batch on the device.  See ddb_batch_create for the meaning of the arguments."""

    def __init__(self, model: EngineModel, num_graphs: int, protein_pos, protein_v, batch_protein,
                 batch_ligand, ligand_v_aux, bond_index, ligand_atom_mask=None, center_mode: int = 0):
        require_cuda()
        L = _lib.lib()
        self.model = model
        self.device = model.device      # the weight blob lives there; every call of this batch runs under that device
        pp, pv = _host(protein_pos, torch.float32), _host(protein_v, torch.float32)
        bp, bl = _host(batch_protein, torch.int64), _host(batch_ligand, torch.int64)
        aux = _host(ligand_v_aux, torch.float32)
        if bond_index is None:
            bond_index = torch.zeros(2, 0, dtype=torch.int64)
        bi = _host(bond_index, torch.int64)
        mask = None if ligand_atom_mask is None else _host(ligand_atom_mask, torch.uint8)
        if pv.dim() != 2 or pv.size(1) != model.cfg['protein_feature_dim']:
            raise ValueError(f'protein_v must be (n, {model.cfg["protein_feature_dim"]})')
        if aux.numel() != bl.numel() * (model.cfg['ligand_feature_dim'] - model.cfg['num_classes'] - (1 if model.time_emb else 0)):
            raise ValueError('ligand_v_aux has the wrong width')
        self.n_protein, self.n_ligand, self.n_bonds = pp.size(0), bl.numel(), bi.size(1)
        self.num_graphs = num_graphs
        self.C, self.Cb = model.cfg['num_classes'], model.cfg['num_bond_classes']
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.ddb_batch_create(
                C.byref(self._h), model._h, num_graphs, self.n_protein, _ptr(pp), _ptr(pv), _ptr(bp),
                self.n_ligand, _ptr(bl), _ptr(aux), self.n_bonds, _ptr(bi), _ptr(mask), center_mode))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            try:
                _lib.lib().ddb_batch_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- state ---------------------------------------------------------------------------------
    def offset(self) -> torch.Tensor:
        out = torch.empty(self.num_graphs, 3, dtype=torch.float32)
        _lib.check(_lib.lib().ddb_batch_get_offset(self._h, _ptr(out)))
        return out

    @_on_device
    def set_state(self, ligand_pos, ligand_v, bond_type):
        dev = self.device
        if bond_type is None:
            bond_type = torch.zeros(0, dtype=torch.int64)
        self._pos = ligand_pos.detach().to(dev, torch.float32).contiguous()
        self._v = ligand_v.detach().to(dev, torch.int64).contiguous()
        self._b = bond_type.detach().to(dev, torch.int64).contiguous()
        if self._pos.shape != (self.n_ligand, 3) or self._v.numel() != self.n_ligand or self._b.numel() != self.n_bonds:
            raise ValueError('state tensors do not match the batch')
        if self._v.numel() and (int(self._v.max()) >= self.C or int(self._v.min()) < 0):
            raise AssertionError(f'Error: {int(self._v.max())} >= {self.C}')      # transitions.py:66
        if self._b.numel() and (int(self._b.max()) >= self.Cb or int(self._b.min()) < 0):
            raise AssertionError(f'Error: {int(self._b.max())} >= {self.Cb}')
        _lib.check(_lib.lib().ddb_batch_set_state(self._h, _ptr(self._pos), _ptr(self._v), _ptr(self._b), _stream_ptr(self.device)))

    @_on_device
    def get_state(self):
        dev = self.device
        pos = torch.empty(self.n_ligand, 3, device=dev, dtype=torch.float32)
        v = torch.empty(self.n_ligand, device=dev, dtype=torch.int64)
        b = torch.empty(self.n_bonds, device=dev, dtype=torch.int64)
        _lib.check(_lib.lib().ddb_batch_get_state(self._h, _ptr(pos), _ptr(v), _ptr(b), _stream_ptr(self.device)))
        return pos, v, b

    # -- compute -------------------------------------------------------------------------------
    @_on_device
    def forward(self):
        dev = self.device
        pos = torch.empty(self.n_ligand, 3, device=dev, dtype=torch.float32)
        vl = torch.empty(self.n_ligand, self.C, device=dev, dtype=torch.float32)
        bl = torch.empty(self.n_bonds, self.Cb, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().ddb_forward(self._h, _ptr(pos), _ptr(vl), _ptr(bl), _stream_ptr(self.device)))
        return pos, vl, bl

    @_on_device
    @_on_device
    def forward_all(self):
        """forward + the head on the input ligand embedding (return_all=True, decompdiff.py:343-350)."""
        dev = self.device
        pos = torch.empty(self.n_ligand, 3, device=dev, dtype=torch.float32)
        vl = torch.empty(self.n_ligand, self.C, device=dev, dtype=torch.float32)
        v0 = torch.empty(self.n_ligand, self.C, device=dev, dtype=torch.float32)
        bl = torch.empty(self.n_bonds, self.Cb, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().ddb_forward_ex(self._h, _ptr(pos), _ptr(vl), _ptr(bl), _ptr(v0), _stream_ptr(self.device)))
        return pos, vl, bl, v0

    @_on_device
    def set_time_steps(self, time_step):
        """Per-graph time steps of a forward() call (only the 'simple' time embedding reads them)."""
        ts = _host(torch.as_tensor(time_step).reshape(-1), torch.int64)
        if ts.numel() != self.num_graphs:
            raise ValueError('time_step needs one entry per graph')
        _lib.check(_lib.lib().ddb_batch_set_time_steps(self._h, C.c_void_p(ts.data_ptr()), _stream_ptr(self.device)))

    def set_time(self, t_start: int):
        _lib.check(_lib.lib().ddb_batch_set_time(self._h, int(t_start), _stream_ptr(self.device)))

    @_on_device
    def set_guidance(self, armsca=None, clash=None, scale_armsca=False, scale_clash=False):
        """armsca = (ligand_decomp_index, min_d, max_d) | None;  clash = (full_pos, full_batch, sigma, gamma) | None;
        scale_* = the drift's `scale: True` option (gradient x pos_score_coef[t])"""
        L = _lib.lib()
        di = _host(armsca[0], torch.int64) if armsca else None
        fp = _host(clash[0], torch.float32) if clash else None
        fb = _host(clash[1], torch.int64) if clash else None
        _lib.check(L.ddb_batch_set_guidance(
            self._h, 1 if armsca else 0, _ptr(di), float(armsca[1]) if armsca else 0.0, float(armsca[2]) if armsca else 0.0,
            1 if clash else 0, fp.size(0) if clash else 0, _ptr(fp), _ptr(fb),
            float(clash[2]) if clash else 0.0, float(clash[3]) if clash else 0.0))
        _lib.check(L.ddb_batch_set_guidance_scale(self._h, int(bool(scale_armsca)), int(bool(scale_clash))))

    @_on_device
    def reverse_step(self, io):
        _lib.check(_lib.lib().ddb_reverse_step(self._h, C.byref(io), _stream_ptr(self.device)))

    def launch_count(self) -> int:
        return int(_lib.lib().ddb_batch_last_launch_count(self._h))

    @_on_device
    def executed_rows(self):
        """(executed, full) kNN-attention destination rows of the last forward (pruning + first-layer cache)."""
        e, f = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().ddb_batch_executed_rows(self._h, C.byref(e), C.byref(f)))
        return int(e.value), int(f.value)

    def h2d_bytes(self) -> int:
        return int(_lib.lib().ddb_batch_h2d_bytes(self._h))

    def profile(self, enable: bool, reset: bool = False):
        _lib.check(_lib.lib().ddb_batch_profile(self._h, int(enable), int(reset)))

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        """{category: {'ms': total device ms, 'count': launches}} accumulated while profiling was enabled."""
        L = _lib.lib()
        n = L.ddb_profile_num_categories()
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(L.ddb_batch_profile_read(self._h, ms, cnt))
        return {L.ddb_profile_category_name(i).decode(): {'ms': ms[i], 'count': int(cnt[i])} for i in range(n) if cnt[i]}

    @_on_device
    def debug_buffer(self, name: str) -> torch.Tensor:
        """Copy of an internal buffer of the last forward (tests / profiling)."""
        p, r, c = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().ddb_batch_debug_buffer(self._h, name.encode(), C.byref(p), C.byref(r), C.byref(c)))
        dtype = torch.int32 if name in ('nbr', 'deg', 'nlig') else torch.float32
        out = torch.empty(r.value, c.value, device=self.device, dtype=dtype)
        _lib.check(_lib.lib().ddb_copy_device(_ptr(out), p, out.numel() * out.element_size(), _stream_ptr(self.device)))
        return out


class RefineBatch:
    """One merged-order graph batch for the refine-net seam (ddb_refine_batch_create / ddb_refine_forward)."""

    def __init__(self, model: EngineModel, batch, mask_ligand, mask_ligand_atom, bond_index):
        require_cuda()
        L = _lib.lib()
        self.model, self.device = model, model.device
        bt, ml = _host(batch, torch.int64), _host(mask_ligand, torch.uint8)
        mla = None if mask_ligand_atom is None else _host(mask_ligand_atom, torch.uint8)
        if bond_index is None:
            bond_index = torch.zeros(2, 0, dtype=torch.int64)
        bi = _host(bond_index, torch.int64)
        self.n_nodes, self.n_bonds = bt.numel(), bi.size(1)
        num_graphs = int(bt.max()) + 1 if bt.numel() else 1
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.ddb_refine_batch_create(C.byref(self._h), model._h, num_graphs, self.n_nodes, _ptr(bt), _ptr(ml), _ptr(mla),
                                                 self.n_bonds, _ptr(bi)))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h:
            try:
                _lib.lib().ddb_batch_destroy(h)
            except Exception:
                pass
            self._h = None

    @_on_device
    def forward(self, h, x, h_bond, tap_layers: int = 0):
        """-> (h, x, h_bond) after the last layer; with `tap_layers` = num_layers also the per-layer stacks (test seam)."""
        dev = self.device
        h = h.detach().to(dev, torch.float32).contiguous()
        x = x.detach().to(dev, torch.float32).contiguous()
        hb = h_bond.detach().to(dev, torch.float32).contiguous()
        if h.shape != (self.n_nodes, 128) or x.shape != (self.n_nodes, 3) or hb.shape != (self.n_bonds, 128):
            raise ValueError('h / x / h_bond do not match the batch')
        ho, xo, hbo = torch.empty_like(h), torch.empty_like(x), torch.empty_like(hb)
        taps = None
        if tap_layers:
            taps = (torch.empty(tap_layers, self.n_nodes, 128, device=dev), torch.empty(tap_layers, self.n_nodes, 4, device=dev),
                    torch.empty(tap_layers, self.n_bonds, 128, device=dev))
            _lib.check(_lib.lib().ddb_batch_set_layer_tap(self._h, _ptr(taps[0]), _ptr(taps[1]), _ptr(taps[2])))
        _lib.check(_lib.lib().ddb_refine_forward(self._h, _ptr(h), _ptr(x), _ptr(hb), _ptr(ho), _ptr(xo), _ptr(hbo), _stream_ptr(dev)))
        if tap_layers:
            _lib.check(_lib.lib().ddb_batch_set_layer_tap(self._h, None, None, None))
            return ho, xo, hbo, taps
        return ho, xo, hbo
