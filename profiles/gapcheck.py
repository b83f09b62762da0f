"""Graph replay vs eager per-step wall time and per-kernel sum (diagnoses inter-kernel gaps): python profiles/gapcheck.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import decompdiff_b200 as ddb
from decompdiff_b200 import synthetic as syn

model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
model.load_state_dict(syn.synthetic_state_dict(model, seed=0))
kw = syn.make_batch(64, 370, (8, 8), 14, seed=1000)
run = model.begin_sampling(**kw, num_steps=1000, center_pos_mode='protein')
run.advance(3)
def timed(fn, n):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(n); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
g1 = timed(run.advance, 10)
g2 = timed(run.advance, 30)
def eager(n):
    for _ in range(n): run.step_eager()
e1 = timed(eager, 5)
run.eb.profile(True, reset=True); eager(3); run.eb.profile(False)
prof = run.eb.profile_read()
ksum = sum(v['ms'] for v in prof.values()) / 3
print(f'lib={os.environ.get("DDB_LIB_PATH", "default")} mask={os.environ.get("DDB_TC_ATTN", "15")}: graph {g1:.2f} / {g2:.2f} ms per step, eager {e1:.2f} ms per step, sum of kernels {ksum:.2f} ms')
