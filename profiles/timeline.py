"""In-kernel phase timeline of the triplet tensor-core kernel (library built with DDB_TIMELINE=1): python profiles/timeline.py"""
import ctypes
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import decompdiff_b200 as ddb
from decompdiff_b200 import _lib, synthetic as syn

PHASES = {
    'trip': ['loop top -> loads issued', 'wait D2 (angular MMA)', 'tmem_ld D2 + z + row gather', 'features(t+1) + A2 hand-over', 'LN stats',
             'quad barrier', 'normalise + ReLU', 'wait main MMA(t-1)', 'drain D + split + tmem_st', 'A hand-over', 'finish epilogue'],
    'knn': ['loop top', 'loads issued + staging', 'wait D2 (distance MMA)', 'z = gather + staged + D2', 'features(t+1) + A2 hand-over', 'LN stats',
            'quad barrier', 'normalise + ReLU', 'wait main MMA(t-1)', 'drain D + split + tmem_st + hand-over', 'row gather + finish epilogue'],
}
model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
model.load_state_dict(syn.synthetic_state_dict(model, seed=0))
kw = syn.make_batch(64, 370, (8, 8), 14, seed=1000)
run = model.begin_sampling(**kw, num_steps=1000, center_pos_mode='protein')
run.step_eager(); run.step_eager()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 128)()
lib = _lib.lib()
half = (ctypes.c_ulonglong * 64)()
assert lib.ddb_debug_timeline_trip(half) == 0
buf[0:64] = half[:]
assert lib.ddb_debug_timeline_knn(half) == 0
buf[64:128] = half[:]
for kid, (name, fam) in enumerate((('trip k', 'trip'), ('trip v', 'trip'), ('knn k (last launch: position layer)', 'knn'), ('knn v', 'knn'))):
    for w in range(2):
        base = (kid * 2 + w) * 16
        tiles = buf[base + 12]
        tot = sum(buf[base + i] for i in range(11))
        print(f'--- {name}, {"warp 0" if w == 0 else "warp 13"}: {tiles} tiles, {tot / max(tiles, 1):.0f} cycles / tile')
        for i, ph in enumerate(PHASES[fam]):
            print(f'   {buf[base + i] / max(tiles, 1):8.0f}  {ph}')
