"""BASELINE cfg 5: atom-count sweep 100 -> 1000 atoms per complex, batch 32, T=200 - step time and the kNN-edge (EGNN) kernel group's
achieved GB/s on the SURVEY 8(d) algorithmic bytes, as JSON lines.   python profiles/sweep_cfg5.py [--steps 200]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import decompdiff_b200 as ddb
from decompdiff_b200 import synthetic as syn

ap = argparse.ArgumentParser(); ap.add_argument('--steps', type=int, default=200); args = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
    os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else 6650.0
model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
model.load_state_dict(syn.synthetic_state_dict(model, seed=0))
B, n_lig = 32, 30
for n_atoms in (100, 200, 400, 700, 1000):
    kw = syn.make_batch(B, n_atoms - n_lig, (8, 8), 14, seed=500 + n_atoms)
    run = model.begin_sampling(**kw, num_steps=args.steps + 40, center_pos_mode='protein', keep_traj=False)
    run.advance(20)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run.advance(args.steps); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    run.eb.profile(True, reset=True)
    for _ in range(3): run.step_eager()
    run.eb.profile(False)
    prof = run.eb.profile_read()
    N, E = B * n_atoms, B * n_atoms * 32
    egnn_ms = sum(prof[c]['ms'] for c in ('knn_attn_k', 'knn_attn_v_node', 'knn_pos_k', 'knn_pos_v')) / 3
    graph_ms = sum(prof[c]['ms'] for c in ('knn_graph', 'edge_weight')) / 3
    alg = 6 * (E * 528.0 + 2 * N * 528.0 + 8.0 * E)
    print(json.dumps({'workload': f'cfg5: {B} pockets x {n_atoms} atoms ({n_lig} ligand), T={args.steps}', 'atoms_per_complex': n_atoms,
                      'ms_per_step': round(ms, 3), 'molecules_per_s_at_T': round(B / (args.steps * ms * 1e-3), 3),
                      'egnn_ms_per_step': round(egnn_ms, 3), 'egnn_alg_GBps': round(alg / (egnn_ms * 1e-3) / 1e9, 1),
                      'egnn_hbm_frac': round(alg / (egnn_ms * 1e-3) / 1e9 / peak, 4), 'graph_build_ms_per_step': round(graph_ms, 3),
                      'graph_build_GBps': round((N * 16.0 + 2 * E * 4.0) / (graph_ms * 1e-3) / 1e9, 1)}), flush=True)
    del run
