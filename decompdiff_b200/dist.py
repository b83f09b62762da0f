"""Multi-GPU plumbing: pockets shard across ranks, one gather of the sampled molecules at the end.

The reference has no distributed code at all - users launch one process per
`--data_id`.  Complexes never interact during the T-step loop (kNN graphs, scatters and priors are all
per-graph), so the B200 layout is: one process per GPU, contiguous blocks of pockets per rank, NO per-step
communication, and a single gather of `pos (sum n,3) f32`, `v (sum n) i64`, `bond (sum Eb) i64` + the per-molecule counts, packed
into ONE buffer per rank and exchanged in ONE `all_gather` - NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` pockets owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _pack(pos, v, bond, na, nb, cap):
    """One uint8 buffer per rank: header (3 x int64 counts) | atoms-per-molecule | bonds-per-molecule | pos | v | bond, every
    section at a fixed offset given by the capacities `cap` = (molecules, atoms, bonds)."""
    M, A, E = cap
    dev = pos.device
    sections = [torch.tensor([na.numel(), pos.size(0), bond.numel()], dtype=torch.int64, device=dev).view(torch.uint8),
                F.pad(na, (0, M - na.numel())).view(torch.uint8), F.pad(nb, (0, M - nb.numel())).view(torch.uint8),
                F.pad(pos.reshape(-1).float(), (0, 3 * (A - pos.size(0)))).view(torch.uint8),
                F.pad(v.long(), (0, A - v.numel())).view(torch.uint8), F.pad(bond.long(), (0, E - bond.numel())).view(torch.uint8)]
    return torch.cat(sections)


def _unpack(buf, cap):
    M, A, E = cap
    sizes = [24, 8 * M, 8 * M, 12 * A, 8 * A, 8 * E]
    head, na, nb, pos, v, bond = buf.split(sizes)
    m, a, e = head.view(torch.int64).tolist()
    # own storage per tensor (views of one byte buffer under several dtypes cannot be saved / are surprising to callers)
    return (pos.view(torch.float32)[:3 * a].view(a, 3).clone(), v.view(torch.int64)[:a].clone(), bond.view(torch.int64)[:e].clone(),
            na.view(torch.int64)[:m].clone(), nb.view(torch.int64)[:m].clone())


def gather_molecules(result: Dict[str, torch.Tensor], atoms_per_mol: Sequence[int], bonds_per_mol: Sequence[int],
                     group=None, capacity: Optional[Tuple[int, int, int]] = None) -> Dict[str, List[torch.Tensor]]:
    """Gather every rank's sampled molecules; returns per-molecule lists in global pocket order
    (rank 0's pockets first).  Without an initialised process group the local result is split and returned.

    The payload of a rank travels as ONE packed buffer in ONE `all_gather`; when the per-rank capacities
    `capacity = (max molecules, max atoms, max bonds)` are not given (they are known a priori when every rank holds shards of
    the same shape) one small `all_gather` of the three counts precedes it."""
    pos, v, bond = result['pos'], result['v'], result['bond']
    dev = pos.device
    na = torch.tensor(list(atoms_per_mol), dtype=torch.int64, device=dev)
    nb = torch.tensor(list(bonds_per_mol), dtype=torch.int64, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        if capacity is None:
            mine = torch.tensor([na.numel(), pos.size(0), bond.numel()], dtype=torch.int64, device=dev)
            allc = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allc, mine, group=group)
            capacity = tuple(int(x) for x in torch.stack(allc).max(0).values.tolist())
        if na.numel() > capacity[0] or pos.size(0) > capacity[1] or bond.numel() > capacity[2]:
            raise ValueError('capacity smaller than the local payload')
        buf = _pack(pos, v, bond, na, nb, capacity)
        out_bufs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out_bufs, buf, group=group)
        parts = [_unpack(b, capacity) for b in out_bufs]
    else:
        parts = [(pos, v, bond, na, nb)]
    out = {'pos': [], 'v': [], 'bond': []}
    for p, vv, bb, a, b in parts:
        out['pos'] += list(p.split(a.tolist()))
        out['v'] += list(vv.split(a.tolist()))
        out['bond'] += list(bb.split(b.tolist()))
    return out
