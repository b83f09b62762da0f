"""The multi-GPU path on CPU: world_size-2 gloo processes shard pockets and gather variable-length molecules."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from decompdiff_b200.dist import gather_molecules, shard_range


def test_shard_range_partitions():
    for n, w in ((512, 8), (64, 8), (10, 4), (3, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _fake_result(rank):
    g = torch.Generator().manual_seed(100 + rank)
    atoms = [5 + rank, 3, 7 + 2 * rank]                 # ragged molecules, different per rank
    bonds = [a * (a - 1) for a in atoms]
    n, e = sum(atoms), sum(bonds)
    return ({'pos': torch.randn(n, 3, generator=g), 'v': torch.randint(0, 8, (n,), generator=g),
             'bond': torch.randint(0, 5, (e,), generator=g)}, atoms, bonds)


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    res, atoms, bonds = _fake_result(rank)
    out = gather_molecules(res, atoms, bonds)
    torch.save(out, os.path.join(out_dir, f'gathered_{rank}.pt'))
    # capacities known a priori: the single packed all_gather alone
    out2 = gather_molecules(res, atoms, bonds, capacity=(3, 40, 400))
    assert all(torch.equal(a, b) for k in out for a, b in zip(out[k], out2[k]))
    dist.destroy_process_group()


def test_gather_molecules_world2(tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = {'pos': [], 'v': [], 'bond': []}
    for r in range(2):
        res, atoms, bonds = _fake_result(r)
        want['pos'] += list(res['pos'].split(atoms))
        want['v'] += list(res['v'].split(atoms))
        want['bond'] += list(res['bond'].split(bonds))
    for r in range(2):     # every rank ends up with the concatenation of the single-rank shards, in rank order
        got = torch.load(os.path.join(str(tmp_path), f'gathered_{r}.pt'))
        for k in want:
            assert len(got[k]) == 6
            assert all(torch.equal(a, b) for a, b in zip(got[k], want[k])), k


def test_gather_without_process_group_splits_locally():
    res, atoms, bonds = _fake_result(0)
    out = gather_molecules(res, atoms, bonds)
    assert [p.size(0) for p in out['pos']] == atoms and [b.numel() for b in out['bond']] == bonds
