import os
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')
RTOL, ATOL = 1e-4, 1e-5     # the tolerance BASELINE.json's north_star states (fp32)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def tol_ratio(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max |got-ref| / (atol + rtol |ref|)  - <= 1 means inside the stated tolerance."""
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    return float(((got - ref).abs() / (ATOL + RTOL * ref.abs())).max()) if ref.numel() else 0.0


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f'{name}.pt'))


@pytest.fixture(scope='session')
def model_cpu():
    """Product module with the name-keyed synthetic weights (CPU tensors, no engine yet)."""
    import decompdiff_b200 as ddb
    from decompdiff_b200 import synthetic as syn
    m = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
    m.load_state_dict(syn.synthetic_state_dict(m, seed=0))
    return m.eval()


@pytest.fixture(scope='session')
def weights(model_cpu):
    return {k: v.detach().clone() for k, v in model_cpu.state_dict().items()}


@pytest.fixture(scope='session')
def oracle_cfg():
    from decompdiff_b200 import synthetic as syn
    c = dict(syn.DEFAULT_MODEL_CONFIG)
    c['num_classes'] = syn.NUM_CLASSES
    return c
