"""Run-to-run repeatability of the forward pass under different kernel selections: python profiles/determinism_check.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import decompdiff_b200 as ddb
from decompdiff_b200 import synthetic as syn
model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
model.load_state_dict(syn.synthetic_state_dict(model, seed=0))
kw = syn.make_batch(n_pockets=3, n_protein=370, arm_sizes=(8, 8), n_scaffold=14, seed=91)
fk = syn.forward_kwargs(kw, None)
ENV = ('DDB_NO_PRUNE', 'DDB_NO_EW_CACHE', 'DDB_TC_ATTN', 'DDB_GEMM')
def run(env):
    for k in ENV:
        os.environ.pop(k, None)
    os.environ.update(env)
    return model(**fk)
results = {}
for name, env in (('default', {}), ('no_prune', {'DDB_NO_PRUNE': '1'}), ('attn simt', {'DDB_TC_ATTN': '0'}), ('gemm simt', {'DDB_GEMM': 'simt'}),
                  ('all simt', {'DDB_TC_ATTN': '0', 'DDB_GEMM': 'simt'})):
    a = run(env); b = run(env); c = run(env)
    results[name] = a
    print(f'{name:14s} run-to-run max |d|', {k: max(float((b[k] - a[k]).abs().max()), float((c[k] - a[k]).abs().max())) for k in a})
for x, y in (('default', 'no_prune'), ('default', 'all simt')):
    print(x, 'vs', y, {k: float((results[x][k] - results[y][k]).abs().max()) for k in results[x]})
