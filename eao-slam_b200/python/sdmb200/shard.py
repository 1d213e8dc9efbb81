"""Keyframe sharding over ranks.

Pass 1 of a keyframe reads only immutable inputs of its neighbours; pass 2 reads the neighbours'
pass-1 planes (ProbabilityMapping.cc:1202-1249).  So the path shards by keyframe with exactly one
exchange step: every rank owns a contiguous range of keyframes, keeps the INPUT planes of the few
neighbour keyframes outside its range as well (halo), and between the passes pulls the halo
keyframes' (rho, sigma) planes from their owners over NVLink.  No reduction, no other collective.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ShardPlan:
    rank: int
    world: int
    per_rank: int            # owned keyframes per rank
    lo: int                  # first global keyframe held locally (owned or halo)
    hi: int                  # one past the last global keyframe held locally
    own_lo: int              # owned range [own_lo, own_hi)
    own_hi: int
    nbr_local: np.ndarray    # [hi-lo, N] local slots of each local keyframe's neighbours (-1: not held)
    halo_local: np.ndarray   # local slots whose pass-1 planes come from a peer
    halo_rank: np.ndarray    # owner rank of each halo keyframe
    halo_peer_slot: np.ndarray  # slot of that keyframe in the owner's arena

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    @property
    def owned_local(self) -> range:
        return range(self.own_lo - self.lo, self.own_hi - self.lo)


def held_range(nbr_global: np.ndarray, own_lo: int, own_hi: int):
    """Contiguous global range covering the owned keyframes and all their neighbours."""
    nb = nbr_global[own_lo:own_hi]
    return int(min(own_lo, nb.min())), int(max(own_hi, nb.max() + 1))


def balanced_bounds(weights, world: int) -> list:
    """Contiguous partition of the keyframes 0 .. len(weights)-1 into `world` ranges of (nearly) equal weight:
    bounds[r] .. bounds[r + 1] is rank r's range.  The cost of a keyframe is proportional to its candidate pixels times the
    columns each of them scans, and both vary along a trajectory (the 1000-keyframe bench trajectory: 48 k - 95 k candidates
    per keyframe), so equal COUNTS leave the slowest rank ~9 % above the mean on 8 GPUs; every rank gets at least one keyframe."""
    w = np.asarray(weights, np.float64)
    G = w.size
    assert G >= world >= 1
    c = np.concatenate([[0.0], np.cumsum(w)])
    bounds = [0]
    for r in range(1, world):
        target = c[-1] * r / world
        i = int(np.searchsorted(c, target))
        if i > 0 and abs(c[i - 1] - target) <= abs(c[min(i, G)] - target):
            i -= 1
        i = min(max(i, bounds[-1] + 1), G - (world - r))
        bounds.append(i)
    bounds.append(G)
    return bounds


def make_plan(nbr_global: np.ndarray, per_rank: int, rank: int, world: int, bounds=None) -> ShardPlan:
    """nbr_global: [G, N] neighbour lists over the whole trajectory.  Ranges are equal (G = per_rank * world) unless
    `bounds` (world + 1 ascending keyframe indices, e.g. balanced_bounds) gives them."""
    G = nbr_global.shape[0]
    if bounds is None:
        assert G == per_rank * world, (G, per_rank, world)
        bounds = [r * per_rank for r in range(world + 1)]
    bounds = [int(b) for b in bounds]
    assert len(bounds) == world + 1 and bounds[0] == 0 and bounds[-1] == G and all(b1 > b0 for b0, b1 in zip(bounds, bounds[1:]))
    own_lo, own_hi = bounds[rank], bounds[rank + 1]
    lo, hi = held_range(nbr_global, own_lo, own_hi)
    nbr_local = np.full((hi - lo, nbr_global.shape[1]), -1, np.int32)
    for g in range(lo, hi):
        for j, v in enumerate(nbr_global[g]):
            if lo <= v < hi:
                nbr_local[g - lo, j] = v - lo
    halo = [g for g in range(lo, hi) if not (own_lo <= g < own_hi)]
    # a halo keyframe is only needed if an OWNED keyframe lists it
    needed = set(int(v) for v in nbr_global[own_lo:own_hi].reshape(-1))
    halo = [g for g in halo if g in needed]
    h_rank = [int(np.searchsorted(bounds, g, side="right")) - 1 for g in halo]
    h_slot = []
    for g, r in zip(halo, h_rank):
        plo, _ = held_range(nbr_global, bounds[r], bounds[r + 1])
        h_slot.append(g - plo)
    return ShardPlan(rank, world, own_hi - own_lo, lo, hi, own_lo, own_hi, nbr_local,
                     np.asarray([g - lo for g in halo], np.int32), np.asarray(h_rank, np.int32),
                     np.asarray(h_slot, np.int32))
