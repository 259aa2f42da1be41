"""Subgroup sharding across ranks (one process per GPU).

RAMBL solves one StrainCall problem per seed gene and the problems never interact
(scripts/rambl.py:165-194 runs them in a process pool), so multi-GPU is: deal the subgroups to the
ranks, let every rank build and solve its own batch, gather the FASTA text on rank 0.  There is no
collective on the data path; ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests) only
carries the final gather.  Costs differ a lot between subgroups (cost ~ levels x draws x strains), so the
deal is a greedy longest-processing-time assignment on a cheap cost proxy rather than round-robin.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence


def cost_proxy(sg) -> float:
    """Work estimate of a subgroup: every graph level draws <= 40000 times per strain set, and the number
    of levels follows the gene length; reads only matter until the depth cap."""
    return float(len(sg.gene)) * float(min(sum(sg.cn), 40000)) + 1.0


def assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Deterministic LPT assignment: indices of the subgroups each rank takes (same answer on every rank)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    mine: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        mine[r].append(i)
        load[r] += costs[i]
    for m in mine:
        m.sort()
    return mine


def solve_sharded(subgroups: Sequence, rank: int, world: int, solve: Callable[[Sequence], List[str]],
                  gather: Optional[Callable[[object], List[object]]] = None) -> Optional[List[str]]:
    """Solve ``subgroups`` across ``world`` ranks.

    ``solve`` maps a list of subgroups to one FASTA string each (on the GPU box:
    ``StrainCallBatch`` build_graphs + infer + fasta).  ``gather`` collects one python object per rank
    (``torch.distributed.all_gather_object`` wrapped by the caller); with world == 1 it is not needed.
    Rank 0 returns the FASTA strings in the ORIGINAL subgroup order, other ranks return None."""
    mine = assign([cost_proxy(s) for s in subgroups], world)[rank]
    local = solve([subgroups[i] for i in mine])
    assert len(local) == len(mine)
    part = list(zip(mine, local))
    parts = [part] if world == 1 or gather is None else gather(part)
    if rank != 0:
        return None
    out: List[Optional[str]] = [None] * len(subgroups)
    for p in parts:
        for i, txt in p:
            out[i] = txt
    assert all(x is not None for x in out)
    return out  # type: ignore[return-value]
