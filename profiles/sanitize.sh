#!/bin/bash
# compute-sanitizer memcheck + racecheck of the whole hot path on a small batch (run under gpurun): profiles/sanitize.sh
mkdir -p gpurun_out
cat > /tmp/san.py <<'P'
import torch, sys
sys.path.insert(0, '.')
import decompdiff_b200 as ddb
from decompdiff_b200 import synthetic as syn
model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
model.load_state_dict(syn.synthetic_state_dict(model, seed=0))
model.use_cuda_graph = False
kw = syn.make_batch(2, 370, (8, 8), 14, seed=3, n_full_extra=40)      # 2 x 400 atoms (cfg-2 pocket shape)
kw2 = syn.make_batch(2, 60, (12, 12), 16, seed=4)                      # 40-atom ligands: chunked softmax groups
r = model.sample_diffusion(**kw, num_steps=2, center_pos_mode='protein',
                           energy_drift_opt=[{'type': 'armsca_prox', 'min_d': 1.2, 'max_d': 1.9}, {'type': 'clash', 'sigma': 2, 'gamma': 4}])
r2 = model.sample_diffusion(**kw2, num_steps=2, center_pos_mode='protein')
torch.cuda.synchronize()
print('ok', r['pos'].shape, bool(torch.isfinite(r['pos']).all()), bool(torch.isfinite(r2['pos']).all()))
P
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?"; tail -4 gpurun_out/sanitize_$tool.log
done
