"""(head, batch) sharding of the attention path across GPUs -- no collective.

Every (kv-head group, batch) pair is independent in forward and backward (the reference's own
grid axes: src/attention.jl:152, src/attention_bwd.jl:263), so a rank simply owns a contiguous
slab of the batch axis (outermost axis => zero-copy views) and, when B < world size, of the
kv-head axis inside one batch element.  GQA groups stay on one rank so dK/dV reduce locally.
"""
from __future__ import annotations


def shard_slices(B: int, KH: int, rank: int, world: int):
    """Return ``(b_slice, kvh_slice)`` owned by ``rank``; units = B*KH (batch-major)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if B % world == 0:
        per = B // world
        return slice(rank * per, (rank + 1) * per), slice(0, KH)
    if world % B == 0 and KH % (world // B) == 0:
        ranks_per_b = world // B
        b = rank // ranks_per_b
        per = KH // ranks_per_b
        r = rank % ranks_per_b
        return slice(b, b + 1), slice(r * per, (r + 1) * per)
    # uneven: contiguous runs of batch elements, sizes differing by at most one
    lo = (B * rank) // world
    hi = (B * (rank + 1)) // world
    return slice(lo, hi), slice(0, KH)


def shard_attention_inputs(q, k, v, rank: int, world: int):
    """Views of q (B,QH,L,E), k, v (B,KH,L,E) owned by ``rank`` (contiguous when only B is cut)."""
    B, QH = q.shape[0], q.shape[1]
    KH = k.shape[1]
    g = QH // KH
    bs, hs = shard_slices(B, KH, rank, world)
    qs = slice(hs.start * g, hs.stop * g)
    return q[bs, qs], k[bs, hs], v[bs, hs]
