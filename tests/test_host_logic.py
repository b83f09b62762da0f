"""Host-side logic that needs no GPU: state_dict contract, batch collation, synthetic shapes, C-ABI surface,
loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch

import decompdiff_b200 as ddb
from conftest import REPO, load_golden
from decompdiff_b200 import _lib, synthetic as syn
from decompdiff_b200 import build as ddb_build


def test_state_dict_matches_reference_keys(model_cpu):
    gold = load_golden('state_dict_keys')          # {name: shape} of the unmodified reference model
    sd = model_cpu.state_dict()
    assert list(sd.keys()) == list(gold.keys())    # same 616 names, same order
    assert len(sd) == 616
    for k, shape in gold.items():
        assert list(sd[k].shape) == shape, k
    assert sum(p.numel() for p in model_cpu.parameters() if p.requires_grad) == 4975269
    model_cpu.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    with pytest.raises(RuntimeError):
        bad = dict(sd)
        bad.pop('refine_net.base_block.0.lin_node.weight')
        model_cpu.load_state_dict(bad, strict=True)


def test_unsupported_configurations_raise():
    cfg = dict(syn.DEFAULT_MODEL_CONFIG, cutoff_mode='ball')
    with pytest.raises(ValueError):                # uni_transformer_edge.py:358 ('knn', 'hybrid' and the product's 'radius' exist)
        ddb.DecompScorePosNet3D(cfg, 29, 10, 8)
    for extra in (dict(add_prior_node=True), dict(time_emb_dim=8, time_emb_mode='sin'), dict(model_type='uni_o2')):
        with pytest.raises((NotImplementedError, ValueError)):
            ddb.DecompScorePosNet3D(dict(syn.DEFAULT_MODEL_CONFIG, **extra), 29, 10, 8)
    m = ddb.DecompScorePosNet3D(dict(syn.DEFAULT_MODEL_CONFIG, time_emb_dim=1, time_emb_mode='simple'), 29, 10, 8)
    assert m.ligand_atom_emb.in_features == 11      # 'simple' time embedding: one more input column (decompdiff.py:173)
    with pytest.raises(ValueError):
        ddb.get_refine_net('uni_o2', ddb.AttrDict(syn.DEFAULT_MODEL_CONFIG))


def test_batch_collate_offsets():
    gen = torch.Generator().manual_seed(0)
    a = syn.make_pocket(gen, 20, (2, 3), 4)
    b = syn.make_pocket(gen, 15, (1,), 2)
    batch = ddb.Batch.from_data_list([a, b], follow_batch=ddb.FOLLOW_BATCH)
    # ligand_decomp_mask += num_arms + 1 per graph (utils/data.py:439-441) -> indexes rows of ligand_decomp_centers
    assert batch.ligand_decomp_mask.tolist() == a.ligand_decomp_mask.tolist() + (b.ligand_decomp_mask + 3).tolist()
    assert torch.equal(batch.ligand_decomp_centers[batch.ligand_decomp_mask],
                       torch.cat([a.ligand_decomp_centers[a.ligand_decomp_mask], b.ligand_decomp_centers[b.ligand_decomp_mask]]))
    # ligand_fc_bond_index += n_ligand (utils/data.py:444), concatenated along the last dim
    assert batch.ligand_fc_bond_index.shape == (2, 9 * 8 + 3 * 2)
    assert int(batch.ligand_fc_bond_index[:, :72].max()) == 8 and int(batch.ligand_fc_bond_index[:, 72:].min()) == 9
    assert batch.protein_element_batch.tolist() == [0] * 20 + [1] * 15
    assert batch.ligand_decomp_centers_batch.tolist() == [0, 0, 0, 1, 1]
    assert batch.ligand_fc_bond_type_batch.tolist() == [0] * 72 + [1] * 6
    assert batch.num_graphs == 2


def test_fc_bond_index_is_dst_major():
    bi = syn.fc_bond_index(4)                      # utils/transforms.py:331-337
    assert bi[1].tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3]
    assert bi[0].tolist() == [1, 2, 3, 0, 2, 3, 0, 1, 3, 0, 1, 2]


def test_cfg2_synthetic_shapes():
    kw = syn.make_batch(4, 370, (8, 8), 14, seed=0)
    assert kw['protein_pos'].shape == (4 * 370, 3) and kw['protein_v'].shape == (4 * 370, 29)
    assert kw['init_ligand_pos'].shape == (120, 3) and kw['ligand_v_aux'].shape == (120, 2)
    assert kw['ligand_fc_bond_index'].shape == (2, 4 * 870) and kw['init_ligand_fc_bond_type'].shape == (4 * 870,)
    assert kw['prior_stds'].shape == (12, 3) and float(kw['prior_stds'].min()) >= 0.6
    assert int(kw['ligand_decomp_batch'].max()) == 11 and int(kw['init_ligand_v'].max()) <= 7
    sd1 = syn.synthetic_state_dict(ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, 29, 10, 8), 0)
    sd2 = syn.synthetic_state_dict(ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, 29, 10, 8), 0)
    assert all(torch.equal(sd1[k], sd2[k]) for k in sd1)     # independent of construction-time RNG


@pytest.fixture(scope='module')
def lib_path():
    return ddb_build.build()


def test_cabi_exports_every_declared_symbol(lib_path):
    header = open(os.path.join(REPO, 'include', 'decompdiff_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(ddb_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(lib_path)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared           # the ctypes binding covers exactly the header
    lib.ddb_version.restype = ctypes.c_char_p
    assert b'sm_100a' in lib.ddb_version()


def test_cabi_argument_errors_without_gpu(lib_path):
    L = _lib.lib()
    h = ctypes.c_void_p()
    cfg = _lib.Config(hidden_dim=64, n_heads=16, knn=32, num_layers=6, num_blocks=1, num_classes=8, num_bond_classes=5,
                      protein_feature_dim=29, ligand_feature_dim=10, num_timesteps=1000)
    assert L.ddb_model_create(ctypes.byref(h), ctypes.byref(cfg)) == 1          # DDB_ERR_INVALID
    assert b'hidden_dim' in L.ddb_last_error()
    with pytest.raises(ValueError):
        _lib.check(1)
    assert L.ddb_forward(None, None, None, None, None) == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the CPU-only behaviour')
def test_product_path_fails_loudly_without_cuda(model_cpu):
    kw = syn.make_batch(1, 20, (2,), 2, seed=0)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model_cpu(**syn.forward_kwargs(kw, None))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model_cpu.sample_diffusion(**kw, num_steps=1, center_pos_mode='protein')
    with pytest.raises(NotImplementedError):
        model_cpu.get_diffusion_loss()


def test_replicated_collate_equals_per_sample_collate():
    """`Batch.from_replicas` (one shared sample + per-sample overrides, assembled with tensor ops) == `Batch.from_data_list`."""
    gen = torch.Generator().manual_seed(7)
    proto = syn.make_pocket(gen, 40, (3, 4), 5)
    n = 4
    bt = [torch.randint(0, 5, (proto.ligand_fc_bond_index.size(1),), generator=gen) for _ in range(n)]
    samples = []
    for k in range(n):
        d = proto.clone()
        d.ligand_fc_bond_type = bt[k]
        samples.append(d)
    a = ddb.Batch.from_data_list(samples, follow_batch=ddb.FOLLOW_BATCH)
    b = ddb.Batch.from_replicas(proto, n, follow_batch=ddb.FOLLOW_BATCH, overrides={'ligand_fc_bond_type': bt})
    assert set(a.keys) == set(b.keys)
    for k in a.keys:
        if torch.is_tensor(a[k]):
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k
