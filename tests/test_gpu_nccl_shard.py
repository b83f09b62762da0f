"""Multi-GPU parity: pockets sharded over 2 ranks (one process per GPU, NCCL) with injected noise must reproduce
the single-GPU run of the whole batch - bit for bit in the discrete samples, within tolerance in the positions - after the one
gather of the path.  Needs 2 CUDA devices (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["DDB_REPO"])
import torch.distributed as dist
import decompdiff_b200 as ddb
from decompdiff_b200 import synthetic as syn
from decompdiff_b200.dist import gather_molecules, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
model = ddb.DecompScorePosNet3D(syn.DEFAULT_MODEL_CONFIG, syn.PROTEIN_FEATURE_DIM, syn.LIGAND_FEATURE_DIM, syn.NUM_CLASSES)
model.load_state_dict(syn.synthetic_state_dict(model, seed=0)); model.eval()
kw = syn.make_batch(n_pockets=6, n_protein=150, arm_sizes=(5, 4), n_scaffold=7, seed=31, ragged=True)
S = 5
n, Eb = kw["init_ligand_pos"].size(0), kw["init_ligand_fc_bond_type"].numel()
noise = syn.step_noise(n, Eb, S, seed=17)
lo, hi = shard_range(6, rank, world)
sub, rows = syn.select_pockets(kw, list(range(lo, hi)))
sub_noise = [{"u_atom": z["u_atom"][rows["ligand"]], "u_bond": z["u_bond"][rows["bond"]], "eps_pos": z["eps_pos"][rows["ligand"]]} for z in noise]
r = model.sample_diffusion(**sub, num_steps=S, center_pos_mode="protein", noise=sub_noise, keep_traj=False)
atoms = torch.bincount(sub["batch_ligand"]).tolist(); bonds = torch.bincount(sub["batch_ligand_bond"]).tolist()
out = gather_molecules({k: r[k].cuda() for k in ("pos", "v", "bond")}, atoms, bonds)
if rank == 0:
    full = model.sample_diffusion(**kw, num_steps=S, center_pos_mode="protein", noise=noise, keep_traj=False)
    pos = torch.cat([p.cpu() for p in out["pos"]]); v = torch.cat([x.cpu() for x in out["v"]]); b = torch.cat([x.cpu() for x in out["bond"]])
    assert len(out["pos"]) == 6
    assert torch.equal(v, full["v"].cpu()) and torch.equal(b, full["bond"].cpu()), "discrete samples differ"
    err = (pos - full["pos"].cpu()).abs(); tol = 1e-5 + 1e-4 * full["pos"].cpu().abs()
    assert bool((err <= tol).all()), float((err / tol).max())
    print("SHARD_OK max err/tol %.4f" % float((err / tol).max()))
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_sharded_run_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, DDB_REPO=REPO)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), str(script)], capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and 'SHARD_OK' in r.stdout
