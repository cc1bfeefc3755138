"""Shared pieces of the ring-attention tests: a CPU stand-in for the kernel backend (test
infrastructure: fp64 torch restatement of the per-block math, checked against the oracle) and the
per-rank worker used by both the gloo (CPU) and NCCL (GPU) runs."""
import math
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "nnop.jl_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


class StandInBackend:
    """fp64 torch version of the five backend calls (same contract as nnop_b200.ring.CudaBackend)."""
    acc_dtype = torch.float64

    @staticmethod
    def _expand(k, g):
        return k.repeat_interleave(g, dim=1)

    def _scores(self, q, k, causal):
        g = q.shape[1] // k.shape[1]
        s = torch.einsum("bhqe,bhke->bhqk", q.double(), self._expand(k.double(), g)) / math.sqrt(q.shape[-1])
        if causal:
            QL, KL = s.shape[-2:]
            keep = torch.arange(KL)[None, :] <= torch.arange(QL)[:, None]   # top-left aligned
            s = s.masked_fill(~keep, -math.inf)
        return s, g

    def attn_fwd(self, q, k, v, causal):
        s, g = self._scores(q, k, causal)
        lse = torch.logsumexp(s, dim=-1)
        p = torch.exp(s - lse[..., None])
        return torch.einsum("bhqk,bhke->bhqe", p, self._expand(v.double(), g)).to(q.dtype), lse.float()

    def attn_bwd(self, dO, o, lse, q, k, v, causal):
        s, g = self._scores(q, k, causal)
        p = torch.exp(s - lse.double()[..., None])          # normalised by the GLOBAL lse
        dO, vv, kk = dO.double(), self._expand(v.double(), g), self._expand(k.double(), g)
        delta = (dO * o.double()).sum(-1, keepdim=True)
        dv = torch.einsum("bhqk,bhqe->bhke", p, dO)
        ds = p * (torch.einsum("bhqe,bhke->bhqk", dO, vv) - delta) / math.sqrt(q.shape[-1])
        dq = torch.einsum("bhqk,bhke->bhqe", ds, kk)
        dk = torch.einsum("bhqk,bhqe->bhke", ds, q.double())
        fold = lambda t: t.reshape(t.shape[0], k.shape[1], g, *t.shape[2:]).sum(2)
        return dq.to(q.dtype), fold(dk).to(q.dtype), fold(dv).to(q.dtype)

    def merge(self, o_acc, lse_acc, o_part, lse_part, init):
        if init:
            o_acc.copy_(o_part)
            lse_acc.copy_(lse_part)
            return lse_acc
        m = torch.maximum(lse_acc, lse_part)
        wa, wp = torch.exp(lse_acc - m), torch.exp(lse_part - m)
        o_acc.copy_((o_acc * wa[..., None] + o_part.to(o_acc.dtype) * wp[..., None]) / (wa + wp)[..., None])
        return m + torch.log(wa + wp)

    def accumulate(self, acc, part, init):
        acc.copy_(part) if init else acc.add_(part.to(acc.dtype))

    def store_rows(self, out, acc, row_offset):
        out[..., row_offset:row_offset + acc.shape[-2], :] = acc.to(out.dtype)


def full_inputs(B, QH, KH, L, E, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, QH, L, E, generator=g).to(dtype)
    k = torch.randn(B, KH, L, E, generator=g).to(dtype)
    v = torch.randn(B, KH, L, E, generator=g).to(dtype)
    dO = torch.randn(B, QH, L, E, generator=g).to(dtype)
    return q, k, v, dO


def ring_worker(rank, world, port, tmp, backend_name, shape, dtype, causal):
    """One rank: shard the (seeded, identical on every rank) inputs, run ring fwd + bwd, save."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    cuda = backend_name == "cuda"
    if cuda:
        torch.cuda.set_device(rank)
    dist.init_process_group("nccl" if cuda else "gloo", rank=rank, world_size=world)
    if cuda:
        import nnop_b200.ring as R
        backend = None
    else:
        # CPU: load only the schedule module (the package itself needs the CUDA library)
        import importlib.util, types
        pkg = types.ModuleType("nnop_b200")
        pkg.__path__ = [str(ROOT / "nnop.jl_b200" / "nnop_b200")]
        sys.modules.setdefault("nnop_b200", pkg)
        if "nnop_b200._lib" not in sys.modules:
            try:
                import nnop_b200._lib  # noqa: F401  (works when the .so is built)
                import nnop_b200.ops  # noqa: F401
            except Exception:
                pass
        import nnop_b200.ring as R
        backend = StandInBackend()
    q, k, v, dO = full_inputs(*shape, dtype)
    shard = (lambda t: R.zigzag_shard(t, rank, world)) if causal else (lambda t: R.contiguous_shard(t, rank, world))
    dev = (lambda t: t.cuda()) if cuda else (lambda t: t)
    ql, kl, vl, dOl = (dev(shard(t)) for t in (q, k, v, dO))
    o, res = R.ring_attention_forward(ql, kl, vl, causal=causal, backend=backend)
    dq, dk, dv = R.ring_attention_backward(dOl, res, causal=causal, backend=backend)
    if cuda:
        torch.cuda.synchronize()
    torch.save(dict(o=o.cpu(), dq=dq.cpu(), dk=dk.cpu(), dv=dv.cpu()), os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def gather(tmp, world, causal):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ring_only", ROOT / "nnop.jl_b200" / "nnop_b200" / "ring.py",
                                                  submodule_search_locations=None)
    parts = [torch.load(os.path.join(tmp, f"r{r}.pt")) for r in range(world)]
    out = {}
    for key in ("o", "dq", "dk", "dv"):
        ps = [p[key] for p in parts]
        if causal:
            c = ps[0].shape[2] // 2
            chunks = [None] * (2 * world)
            for r, p in enumerate(ps):
                chunks[r] = p[:, :, :c]
                chunks[2 * world - 1 - r] = p[:, :, c:]
            out[key] = torch.cat(chunks, dim=2)
        else:
            out[key] = torch.cat(ps, dim=2)
    return out
