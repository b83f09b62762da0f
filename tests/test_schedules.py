"""Schedule known-answers extracted from the reference pin both the product tables
(decompdiff_b200/schedules.py) and the oracle's (oracle/restate.py)."""
import numpy as np
import pytest
import torch

from decompdiff_b200 import schedules, synthetic as syn
from decompdiff_b200.decompdiff import AttrDict
from oracle import restate

# quantity -> values at t = 0, 1, 500, 999   (/root/reference/models/decompdiff.py:96-131, transitions.py:103-113)
KNOWN = {
    'betas': [5.044998943e-06, 5.104607226e-06, 1.003052806e-03, 1.995055005e-03],
    'alphas_cumprod': [9.999949336e-01, 9.999898672e-01, 8.904048800e-01, 3.675538898e-01],
    'posterior_mean_c0_coef': [1.0, 5.029364824e-01, 8.640606888e-03, 1.914368360e-03],
    'posterior_mean_ct_coef': [0.0, 4.970635176e-01, 9.913449883e-01, 9.978413582e-01],
    'posterior_logvar': [-1.288440228e+01, -1.288440228e+01, -6.912898064e+00, -6.218245983e+00],
    'atom_type_trans.log_alphas_cumprod_v': [-2.539948946e-05, -5.321847129e-05, -3.558717370e-01, -9.919879913e+00],
    'atom_type_trans.log_one_minus_alphas_cumprod_v': [-1.058079433e+01, -9.841131210e+00, -1.205849528e+00, -4.918826744e-05],
}
IDX = [0, 1, 500, 999]


def _product_tables():
    cfg = AttrDict(syn.DEFAULT_MODEL_CONFIG)
    out = {k: torch.from_numpy(v) for k, v in schedules.position_tables(cfg).items()}
    for name, K in (('atom_type_trans', 8), ('bond_type_trans', 5)):
        for k, v in schedules.categorical_tables(1000, cfg.v_beta_s, K).items():
            out[f'{name}.{k}'] = torch.from_numpy(v)
    return out


@pytest.mark.parametrize('source', ['product', 'oracle'])
def test_known_answers(source):
    tab = _product_tables() if source == 'product' else restate.schedule_tables(dict(syn.DEFAULT_MODEL_CONFIG, num_classes=8))
    for key, want in KNOWN.items():
        got = tab[key][IDX].double().numpy()
        np.testing.assert_allclose(got, np.array(want), rtol=2e-6, atol=1e-9, err_msg=key)


def test_product_and_oracle_tables_identical():
    a, b = _product_tables(), restate.schedule_tables(dict(syn.DEFAULT_MODEL_CONFIG, num_classes=8))
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == torch.float32 and torch.equal(a[k], b[k]), k


def test_uniform_priors_and_constants():
    tab = _product_tables()
    np.testing.assert_allclose(tab['atom_type_trans.prior_probs'].numpy(), -np.log(8) * np.ones((1, 8)), rtol=1e-6)
    np.testing.assert_allclose(tab['bond_type_trans.prior_probs'].numpy(), -np.log(5) * np.ones((1, 5)), rtol=1e-6)
    assert restate.GAUSS_OFFSETS == [0, 1, 1.25, 1.5, 1.75, 2, 2.25, 2.5, 2.75, 3, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10]
    np.testing.assert_allclose(restate.SSP_SHIFT, np.log(2.0))


def test_other_beta_schedules():
    for kind in ('linear', 'quad', 'const', 'jsd'):
        b = schedules.beta_schedule(kind, 1e-4, 2e-2, 100)
        assert b.shape == (100,) and np.all(b > 0)
    with pytest.raises(NotImplementedError):
        schedules.beta_schedule('nope', 1e-4, 2e-2, 10)
