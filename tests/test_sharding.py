"""(head, batch) sharding logic: coverage/disjointness, and a world_size-2 gloo run on CPU in
which each rank evaluates its shard (with the oracle standing in for the kernels) and the
gathered result equals the unsharded one -- i.e. the path needs no collective."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200" / "nnop_b200"))


def _sharding():
    import importlib.util
    spec = importlib.util.spec_from_file_location("nnop_sharding", ROOT / "nnop.jl_b200" / "nnop_b200" / "sharding.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("B,KH,world", [(8, 32, 1), (8, 32, 2), (8, 8, 4), (8, 8, 8), (1, 8, 8),
                                        (2, 8, 8), (3, 4, 2), (5, 2, 4)])
def test_shards_partition_all_units(B, KH, world):
    sh = _sharding()
    seen = torch.zeros(B, KH, dtype=torch.int32)
    for r in range(world):
        bs, hs = sh.shard_slices(B, KH, r, world)
        seen[bs, hs] += 1
    assert (seen == 1).all()


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    from oracle import oracle as O
    sh = _sharding()
    g = torch.Generator().manual_seed(0)
    B, QH, KH, L, E = 2, 4, 2, 24, 16
    q = torch.randn(B, QH, L, E, generator=g, dtype=torch.float64)
    k = torch.randn(B, KH, L, E, generator=g, dtype=torch.float64)
    v = torch.randn(B, KH, L, E, generator=g, dtype=torch.float64)
    dO = torch.randn(B, QH, L, E, generator=g, dtype=torch.float64)
    qs, ks, vs = sh.shard_attention_inputs(q, k, v, rank, world)
    dOs, _, _ = sh.shard_attention_inputs(dO, k, v, rank, world)
    o = O.naive_attention(qs, ks, vs, causal=True)
    dq, dk, dv, _ = O.naive_attention_bwd(dOs, qs, ks, vs, causal=True)
    # the only cross-rank step is the max-over-ranks timing reduction bench.py does
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    torch.save(dict(o=o, dq=dq, dk=dk, dv=dv), os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_reproduce_full_result(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    g = torch.Generator().manual_seed(0)
    B, QH, KH, L, E = 2, 4, 2, 24, 16
    q = torch.randn(B, QH, L, E, generator=g, dtype=torch.float64)
    k = torch.randn(B, KH, L, E, generator=g, dtype=torch.float64)
    v = torch.randn(B, KH, L, E, generator=g, dtype=torch.float64)
    dO = torch.randn(B, QH, L, E, generator=g, dtype=torch.float64)
    o = O.naive_attention(q, k, v, causal=True)
    dq, dk, dv, _ = O.naive_attention_bwd(dO, q, k, v, causal=True)
    parts = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert torch.equal(torch.cat([p["o"] for p in parts]), o)
    assert torch.equal(torch.cat([p["dq"] for p in parts]), dq)
    assert torch.equal(torch.cat([p["dk"] for p in parts]), dk)
    assert torch.equal(torch.cat([p["dv"] for p in parts]), dv)


def test_bench_pinned_buffer_affinity_helper_restores_affinity():
    """bench.py allocates its pinned host buffers from the CPUs NVML reports as local to the GPU; whatever
    NVML says (here: no driver at all), the helper must leave the process affinity as it found it."""
    import importlib.util
    import os
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    with bench.gpu_local_cpus(0) as note:
        assert isinstance(note[0], str) and note[0]
        assert os.sched_getaffinity(0) <= before
    assert os.sched_getaffinity(0) == before
