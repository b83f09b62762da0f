"""Where do the triplet kernels deviate?  Layer-0 h_bond of the refine net: tensor-core kernels vs the SIMT kernels (DDB_TC_ATTN=28)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from decompdiff_b200 import synthetic as syn
from decompdiff_b200.decompdiff import as_config, get_refine_net
from decompdiff_b200.engine import RefineBatch
import decompdiff_b200 as ddb
from test_gpu_refine_seam import _merged_inputs

model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
weights = syn.synthetic_state_dict(model, seed=0)
kw = syn.make_batch(n_pockets=1, n_protein=120, arm_sizes=(8, 8), n_scaffold=14, seed=81)
h, x, bond_index, h_bond, mask_l, mask_la, batch = _merged_inputs(weights, kw)
net = get_refine_net('uni_o2_bond', as_config(syn.DEFAULT_MODEL_CONFIG))
net.load_state_dict({k[len('refine_net.'):]: v for k, v in weights.items() if k.startswith('refine_net.')})
taps = {}
for mode in ('31', '28'):
    os.environ['DDB_TC_ATTN'] = mode
    rb = RefineBatch(net._refine_engine(torch.device('cuda', 0)), batch, mask_l, mask_la, bond_index)
    _, _, _, (th, tx, thb) = rb.forward(h, x, h_bond, tap_layers=6)
    taps[mode] = thb.cpu()
a, b = taps['31'][0], taps['28'][0]
print('nan in tc', int(torch.isnan(a).sum()), 'nan in simt', int(torch.isnan(b).sum()))
err = torch.nan_to_num((a - b).abs(), nan=1e9)
print('edges', a.shape, 'max err', float(err.max()), 'mean', float(err.mean()))
per_edge = err.max(dim=1).values
bad = (per_edge > 1e-3).nonzero().flatten()
print('bad edges', bad.numel(), 'of', a.shape[0], bad[:64].tolist())
per_ch = err.max(dim=0).values
print('bad channels', (per_ch > 1e-3).nonzero().flatten().tolist()[:128])
src, dst = bond_index[0].tolist(), bond_index[1].tolist()
order = sorted(range(a.shape[0]), key=lambda e: (src[e], dst[e]))
pos_of = {e: p for p, e in enumerate(order)}
n = a.shape[0]
grid = min(148, (n + 3) // 4)
per = (n + 4 * grid - 1) // (4 * grid)
print('n_groups', n, 'grid', grid, 'per', per)
from collections import Counter
print('bad by iteration', Counter(pos_of[int(e)] % per for e in bad), 'bad by quadrant', Counter((pos_of[int(e)] // per) % 4 for e in bad))
print('bad positions', sorted(pos_of[int(e)] for e in bad)[:80])
if bad.numel():
    e = int(bad[0]); print('edge', e, 'tc', a[e, :8].tolist(), 'simt', b[e, :8].tolist(), 'in', h_bond[e, :8].tolist())
