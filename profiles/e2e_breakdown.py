"""Where the end-to-end seconds of sample_diffusion(T=1000, cfg2) go besides the step loop: python profiles/e2e_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import decompdiff_b200 as ddb
from decompdiff_b200 import synthetic as syn
model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
model.load_state_dict(syn.synthetic_state_dict(model, seed=0))
kw = syn.make_batch(64, 370, (8, 8), 14, seed=1000)
host_kw = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in kw.items()}
model.sample_diffusion(**host_kw, num_steps=8, center_pos_mode='protein')      # warm: engine, kernels
def t():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = t(); run = model.begin_sampling(**host_kw, num_steps=1000, center_pos_mode='protein'); t1 = t()
    run.enable_host_streaming(); t2 = t()
    run.advance(1000); t3 = t()
    r = run.finish(); t4 = t()
    print(f'rep {rep}: begin_sampling {t1 - t0:.3f} s, pinned alloc {t2 - t1:.3f} s, 1000 steps (+streaming) {t3 - t2:.3f} s, finish {t4 - t3:.3f} s, total {t4 - t0:.3f} s')
    del run, r
