"""ctypes binding of include/decompdiff_b200.h.

The product path has no CPU fallback: if the shared library cannot be loaded every entry point
raises `RuntimeError` (build it with `python -m decompdiff_b200.build` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libdecompdiff_b200.so')

# every symbol declared in include/decompdiff_b200.h
SYMBOLS = [
    'ddb_last_error', 'ddb_version',
    'ddb_model_create', 'ddb_model_set_tensor', 'ddb_model_finalize', 'ddb_model_destroy',
    'ddb_batch_create', 'ddb_batch_destroy', 'ddb_batch_get_offset', 'ddb_batch_set_state', 'ddb_batch_get_state',
    'ddb_forward', 'ddb_batch_set_time', 'ddb_reverse_step', 'ddb_batch_set_guidance',
    'ddb_knn_graph', 'ddb_gemm128', 'ddb_batch_debug_buffer', 'ddb_copy_device', 'ddb_batch_last_launch_count', 'ddb_batch_h2d_bytes',
    'ddb_batch_profile', 'ddb_profile_num_categories', 'ddb_profile_category_name', 'ddb_batch_profile_read',
    'ddb_batch_executed_rows', 'ddb_model_set_refine_only', 'ddb_forward_ex', 'ddb_refine_batch_create', 'ddb_refine_forward', 'ddb_batch_set_layer_tap', 'ddb_batch_set_guidance_scale', 'ddb_model_set_cutoff', 'ddb_model_set_mean_type', 'ddb_model_set_time_emb', 'ddb_batch_set_time_steps',
]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'hidden_dim', 'n_heads', 'knn', 'num_layers', 'num_blocks', 'num_classes', 'num_bond_classes',
        'protein_feature_dim', 'ligand_feature_dim', 'num_timesteps')]


class StepIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'prior_std_atom', 'u_atom', 'u_bond', 'eps_pos',
        'pos_traj', 'v_traj', 'v0_traj', 'vt_traj', 'bond_traj', 'bt_traj')]


_lib = None


def lib():
    """Load (once) and return the CUDA library; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get('DDB_LIB_PATH', LIB_PATH)      # A/B testing of kernel variants; the default is the in-tree build
    if not os.path.exists(path):
        raise RuntimeError(
            f'{path} is missing - the CUDA extension is required (no CPU fallback). '
            'Build it with `python -m decompdiff_b200.build`.')
    L = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.ddb_last_error.restype = C.c_char_p
    L.ddb_version.restype = C.c_char_p
    L.ddb_model_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.ddb_model_set_tensor.argtypes = [vp, C.c_char_p, vp, i64]
    L.ddb_model_finalize.argtypes = [vp]
    L.ddb_model_destroy.argtypes = [vp]
    L.ddb_model_destroy.restype = None
    L.ddb_batch_create.argtypes = [C.POINTER(vp), vp, i32, i64, vp, vp, vp, i64, vp, vp, i64, vp, vp, i32]
    L.ddb_batch_destroy.argtypes = [vp]
    L.ddb_batch_destroy.restype = None
    L.ddb_batch_get_offset.argtypes = [vp, vp]
    L.ddb_batch_set_state.argtypes = [vp, vp, vp, vp, vp]
    L.ddb_batch_get_state.argtypes = [vp, vp, vp, vp, vp]
    L.ddb_forward.argtypes = [vp, vp, vp, vp, vp]
    L.ddb_batch_set_time.argtypes = [vp, i32, vp]
    L.ddb_reverse_step.argtypes = [vp, C.POINTER(StepIO), vp]
    L.ddb_batch_set_guidance.argtypes = [vp, i32, vp, f32, f32, i32, i64, vp, vp, f32, f32]
    L.ddb_knn_graph.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.ddb_gemm128.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp]
    L.ddb_batch_debug_buffer.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]
    L.ddb_copy_device.argtypes = [vp, vp, i64, vp]
    L.ddb_batch_profile.argtypes = [vp, i32, i32]
    L.ddb_profile_num_categories.restype = i32
    L.ddb_profile_category_name.argtypes = [i32]
    L.ddb_profile_category_name.restype = C.c_char_p
    L.ddb_batch_profile_read.argtypes = [vp, vp, vp]
    L.ddb_model_set_refine_only.argtypes = [vp, i32]
    L.ddb_forward_ex.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ddb_refine_batch_create.argtypes = [C.POINTER(vp), vp, i32, i64, vp, vp, vp, i64, vp]
    L.ddb_refine_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.ddb_model_set_cutoff.argtypes = [vp, i32, f32]
    L.ddb_model_set_mean_type.argtypes = [vp, i32]
    L.ddb_model_set_time_emb.argtypes = [vp, i32]
    L.ddb_batch_set_time_steps.argtypes = [vp, vp, vp]
    L.ddb_batch_set_guidance_scale.argtypes = [vp, i32, i32]
    L.ddb_batch_set_layer_tap.argtypes = [vp, vp, vp, vp]
    L.ddb_batch_executed_rows.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.ddb_batch_last_launch_count.argtypes = [vp]
    L.ddb_batch_last_launch_count.restype = i64
    L.ddb_batch_h2d_bytes.argtypes = [vp]
    L.ddb_batch_h2d_bytes.restype = i64
    for s in SYMBOLS:
        if s not in ('ddb_last_error', 'ddb_version', 'ddb_model_destroy', 'ddb_batch_destroy',
                     'ddb_batch_last_launch_count', 'ddb_batch_h2d_bytes', 'ddb_profile_num_categories',
                     'ddb_profile_category_name'):
            getattr(L, s).restype = C.c_int
    _lib = L
    return L


def check(status: int):
    """Map a ddb_status onto the exception the reference would raise for the same condition."""
    if status == 0:
        return
    msg = lib().ddb_last_error().decode()
    if status == 1:
        raise ValueError(msg)          # reference: ValueError on unsupported modes (decompdiff.py:610,674)
    if status == 2:
        raise KeyError(msg)            # load_state_dict(strict=True) missing key
    raise RuntimeError(f'decompdiff_b200 status {status}: {msg}')
