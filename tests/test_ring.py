"""Sequence-sharded ring attention (BASELINE config 5).

CPU (gloo, world_size 2 and 3): the ring schedule, K/V rotation, log-sum-exp merge order and the
travelling dK/dV accumulators, with an fp64 stand-in for the kernels -- the gathered result must
equal the oracle's attention over the FULL sequence.  GPU (NCCL, 2 ranks; `-m gpu`, needs 2 GPUs):
the same run through libnnop_b200.so."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from helpers import kernel_err, max_abs
from oracle import oracle as O
import ring_common as RC


def _run(tmp_path, world, backend, shape, dtype, causal):
    port = 29600 + (os.getpid() % 1500) + world
    mp.spawn(RC.ring_worker, args=(world, port, str(tmp_path), backend, shape, dtype, causal),
             nprocs=world, join=True)
    got = RC.gather(str(tmp_path), world, causal)
    q, k, v, dO = RC.full_inputs(*shape, dtype)
    D = lambda t: t.double()
    ro = O.naive_attention(D(q), D(k), D(v), causal=causal)
    rq, rk, rv, _ = O.naive_attention_bwd(D(dO), D(q), D(k), D(v), causal=causal)
    return got, dict(o=ro, dq=rq, dk=rk, dv=rv)


def test_stand_in_backend_matches_oracle():
    """The CPU stand-in's per-block fwd/bwd (with residuals) equals the oracle on a whole problem."""
    be = RC.StandInBackend()
    q, k, v, dO = RC.full_inputs(2, 4, 2, 24, 16, torch.float64)
    for causal in (False, True):
        o, lse = be.attn_fwd(q, k, v, causal)
        ro, rl = O.naive_attention(q, k, v, causal=causal, return_lse=True)
        assert max_abs(o, ro) < 1e-12 and max_abs(lse, rl) < 1e-5
        dq, dk, dv = be.attn_bwd(dO, ro, rl, q, k, v, causal)
        rq, rk, rv, _ = O.naive_attention_bwd(dO, q, k, v, causal=causal)
        assert max(max_abs(dq, rq), max_abs(dk, rk), max_abs(dv, rv)) < 1e-10


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_schedule_covers_every_visible_block_pair_once(world):
    """Union over ranks and steps of (global q chunk, global kv chunk, masked?) = the causal
    lower-triangular block structure, each pair exactly once, equal work per rank and step."""
    import importlib.util, sys, types
    pkg = types.ModuleType("nnop_b200"); pkg.__path__ = [str(RC.ROOT / "nnop.jl_b200" / "nnop_b200")]
    sys.modules.setdefault("nnop_b200", pkg)
    try:
        import nnop_b200.ring as R
    except Exception as e:  # the package needs its CUDA library; build it first
        pytest.skip(f"nnop_b200 not importable: {e}")
    seen = {}
    for r in range(world):
        own = R.zigzag_chunks(r, world)
        for src, pairs in R.ring_schedule(r, world, True):
            theirs = R.zigzag_chunks(src, world)
            work = sum(0.5 if cz else 1.0 for _, _, cz in pairs)
            assert work == 2.0
            for qc, kc, cz in pairs:
                key = (own[qc], theirs[kc])
                assert key not in seen
                seen[key] = cz
    n = 2 * world
    for a in range(n):
        for b in range(n):
            if b < a:
                assert seen.get((a, b)) is False
            elif b == a:
                assert seen.get((a, b)) is True
            else:
                assert (a, b) not in seen


@pytest.mark.parametrize("world,causal", [(2, True), (2, False), (3, True)])
def test_ring_gloo_matches_full_attention(tmp_path, world, causal):
    shape = (1, 4, 2, 12 * world, 16)
    got, ref = _run(tmp_path, world, "standin", shape, torch.float64, causal)
    for key in ref:
        assert max_abs(got[key], ref[key]) < 5e-6, key  # lse travels as Float32


@pytest.mark.gpu
@pytest.mark.parametrize("causal", [True, False])
def test_ring_nccl_two_gpus(tmp_path, causal):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    shape = (1, 4, 2, 1024, 128)
    got, ref = _run(tmp_path, 2, "cuda", shape, torch.bfloat16, causal)
    for key in ref:
        assert kernel_err(got[key], ref[key]) < 2e-2, key
