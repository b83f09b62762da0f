"""decompdiff_b200 - B200-native (sm_100a) implementation of DecompDiff's diffusion sampling hot path.

Public surface (mirrors the reference's Python API for the path):
    DecompScorePosNet3D            models/decompdiff.py:75
    get_refine_net                 models/encoders/__init__.py:5
    Batch / ProteinLigandData      torch_geometric Batch + utils/data.py:367 (PyG-free stand-ins)
    FOLLOW_BATCH                   datasets/pl_data.py:11
    sample_diffusion_ligand_decomp scripts/sample_diffusion_decomp.py:57      (sampling.py)
    transforms.*                   utils/transforms.py:114-391                (sampling-time transforms)
    prior.*                        utils/prior.py:11-159                      (decomposed priors)
"""
from .batch import Batch, Data, FOLLOW_BATCH, ProteinLigandData  # noqa: F401
from .decompdiff import AttrDict, DecompScorePosNet3D, get_refine_net  # noqa: F401
from . import prior, transforms  # noqa: F401
from .sampling import log_sample_categorical, sample_diffusion_ligand_decomp, save_results  # noqa: F401

__version__ = '0.1.0'
