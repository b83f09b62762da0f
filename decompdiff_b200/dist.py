"""Multi-GPU plumbing: pockets shard across ranks, one gather of the sampled molecules at the end.

The reference has no distributed code at all - users launch one process per
`--data_id`.  Complexes never interact during the T-step loop (kNN graphs, scatters and priors are all
per-graph), so the B200 layout is: one process per GPU, contiguous blocks of pockets per rank, NO per-step
communication, and a single variable-length gather (`all_gather` of counts, then of padded payloads) of
`pos (sum n,3) f32`, `v (sum n) i64`, `bond (sum Eb) i64` - NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` pockets owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _gather_varlen(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    """all_gather of tensors whose first dimension differs per rank (padded to the max, then trimmed)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.size(0)], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes) if sizes else 0
    buf = t.new_zeros((pad,) + tuple(t.shape[1:]))
    buf[:t.size(0)] = t
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def gather_molecules(result: Dict[str, torch.Tensor], atoms_per_mol: Sequence[int], bonds_per_mol: Sequence[int],
                     group=None) -> Dict[str, List[torch.Tensor]]:
    """Gather every rank's sampled molecules; returns per-molecule lists in global pocket order
    (rank 0's pockets first).  Without an initialised process group the local result is split and returned."""
    pos, v, bond = result['pos'], result['v'], result['bond']
    dev = pos.device
    na = torch.tensor(list(atoms_per_mol), dtype=torch.int64, device=dev)
    nb = torch.tensor(list(bonds_per_mol), dtype=torch.int64, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        pos_l, v_l, bond_l = _gather_varlen(pos, group), _gather_varlen(v, group), _gather_varlen(bond, group)
        na_l, nb_l = _gather_varlen(na, group), _gather_varlen(nb, group)
    else:
        pos_l, v_l, bond_l, na_l, nb_l = [pos], [v], [bond], [na], [nb]
    out = {'pos': [], 'v': [], 'bond': []}
    for p, vv, bb, a, b in zip(pos_l, v_l, bond_l, na_l, nb_l):
        out['pos'] += list(p.split(a.tolist()))
        out['v'] += list(vv.split(a.tolist()))
        out['bond'] += list(bb.split(b.tolist()))
    return out
