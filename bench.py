#!/usr/bin/env python
"""bench.py -- the headline measurement of BASELINE.json:
"flash_attention fwd+bwd TFLOP/s (bf16, E=128, causal)" on config C2
(BF16 causal E=128 L=8192 H=32 B=8), one step = one forward + one backward over the batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, RANK/LOCAL_RANK/WORLD_SIZE from the env).  The
path shards over the independent (head, batch) axis with no data-path collective:
every rank runs its own B=8 slab of a global batch of 8*N ("scaling": "weak"; `--scaling strong`
shards the one B=8 batch instead, 8/N batch elements per rank); NCCL is used only for the barrier
and the max-over-ranks reduction of the device-side time.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same metric through
the host-buffer entry point (pinned host q,k,v,dO in; o,dq,dk,dv out; copies inside the timed
region).  `roofline`: the backward main kernel (dominant launch) against the measured bf16 peak.
`cpu_baseline`: the oracle (reference's naive attention restated in torch-CPU) on a bounded sample.
`--impl reference` times that CPU path alone, on the same config/metric/unit.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "nnop.jl_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

METRIC = "flash_attention fwd+bwd TFLOP/s (bf16, E=128, causal)"
B, H, KH, L, E = 8, 32, 32, 8192, 128
CAUSAL = True


def flops(b=B, h=H, l=L, e=E, causal=CAUSAL):
    f_fwd = 4.0 * b * h * l * l * e * (0.5 if causal else 1.0)   # SURVEY.md §8d convention
    return f_fwd, 2.5 * f_fwd


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 1590.0, 1400.0, "fallback"


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [x for x in vis.split(",") if x.strip() != ""]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # NVML unavailable: report nulls rather than guess
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(sample_b: int, sample_h: int, threads: int):
    """One bounded sample of the reference's CPU path: naive attention forward + closed-form
    backward (oracle/oracle.py, restating test/attention_testsetup.jl:21-45) in Float32 on the host
    cores, for `sample_b x sample_h` (batch, head) units of the C2 shape.  Returns seconds."""
    import torch
    from oracle import oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(sample_b, sample_h, L, E, generator=g)
    k = torch.randn(sample_b, sample_h, L, E, generator=g)
    v = torch.randn(sample_b, sample_h, L, E, generator=g)
    dO = torch.randn(sample_b, sample_h, L, E, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.naive_attention_bwd(dO, q, k, v, causal=CAUSAL)   # recomputes the forward inside
    return time.perf_counter() - t0


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (its naive attention;
    the fused kernels are cpu=false) on this box's host cores.  Rank 0 only."""
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    sb, sh = 1, 4
    for _ in range(args.warmup):
        cpu_reference_step(sb, sh, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_reference_step(sb, sh, threads)
    ms = 1e3 * t / max(1, args.steps)
    f_fwd, f_bwd = flops(sb, sh)
    val = (f_fwd + f_bwd) / (ms * 1e-3) / 1e12
    sample = f"B={sb},H={sh} of the C2 shape (L={L}, E={E}, causal) per step, Float32, torch-CPU naive attention fwd+bwd"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: BF16 causal E=128 L=8192 H=32 B=8 fwd+bwd", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


@contextlib.contextmanager
def gpu_local_cpus(local_rank):
    """Pin this thread to the CPUs NVML reports as local to the GPU while the pinned host buffers are
    allocated (first touch places their pages on that NUMA node), then restore the affinity."""
    note = ["default placement"]
    old = None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 16)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        old = os.sched_getaffinity(0)
        local = cpus & old
        if local and local != old:
            os.sched_setaffinity(0, local)
            note[0] = f"allocated from {len(local)} GPU-local CPUs of {len(old)} (NVML affinity)"
        else:
            old = None
    except Exception as e:  # NVML missing or affinity not permitted: keep the default placement
        old = None
        note[0] = f"default placement ({type(e).__name__})"
    try:
        yield note
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)


# ---------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    import nnop_b200 as nn  # raises ImportError if libnnop_b200.so is missing: no fallback

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    strong = args.scaling == "strong"
    if strong and B % world:
        raise SystemExit(f"--scaling strong shards the batch of {B} over the ranks: {world} does not divide it")
    Bl = B // world if strong else B        # batch elements per rank: (head, batch) sharding, batch first
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    mk = lambda h: torch.randn(Bl, h, L, E, device=dev, dtype=torch.float32, generator=g).to(torch.bfloat16)
    q, k, v, dO = mk(H), mk(KH), mk(KH), mk(H)

    def step():
        o, lse = nn._flash_attention(q, k, v, causal=CAUSAL)
        return nn.grad_flash_attention(dO, o, lse, q, k, v, causal=CAUSAL)

    for _ in range(max(args.warmup, 3)):
        step()
    assert nn.last_attention_path() == 1, "tcgen05 path did not run"
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b_ in kev:   # materialise the cudaEvent handles before handing them to the library
        a.record(); b_.record()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for i in range(args.steps):
            nn.set_timing_events(1, kev[i][0], kev[i][1])  # brackets the backward main kernel only
            step()
        ev1.record()
        torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    kern_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in kev)
    t = torch.tensor([ms_total], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    f_fwd, f_bwd = flops(Bl)                # per rank
    value = world * (f_fwd + f_bwd) / (ms_step * 1e-3) / 1e12

    # ---- end to end through the host-buffer entry point --------------------------------
    e2e = None
    if not args.no_e2e:
        pin = lambda t_: torch.empty(t_.shape, dtype=t_.dtype, pin_memory=True).copy_(t_)
        with gpu_local_cpus(local_rank) as numa_note:   # first touch: pinned pages on the GPU's NUMA node
            hq, hk, hv, hdO = pin(q), pin(k), pin(v), pin(dO)
            out = {n: torch.empty(s.shape, dtype=s.dtype, pin_memory=True) for n, s in
                   (("o", q), ("dq", q), ("dk", k), ("dv", k))}
        pipe = nn.HostAttentionPipeline(q.shape, k.shape, torch.bfloat16, causal=CAUSAL, chunk=1,
                                        kv_heads=args.e2e_kv_heads, device=dev)
        for _ in range(2):
            pipe(hq, hk, hv, hdO, out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            pipe(hq, hk, hv, hdO, out)
        e1.record()
        torch.cuda.synchronize()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e2e = te.item() / args.steps
        e2e = {"value": world * (f_fwd + f_bwd) / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
               "ms_per_step": ms_e2e,
               "pipeline": f"{Bl * (KH // pipe.kv_heads)} chunks of {pipe.kv_heads} kv heads x 1 batch element, "
                           f"{len(pipe.slots)} device slots, H2D / compute / D2H on three streams",
               "pinned_host_memory": numa_note[0]}
        del hq, hk, hv, hdO, out, pipe

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    burst, sustained, src = measured_peaks()
    achieved = f_bwd / (kern_ms * 1e-3) / 1e12
    # dram__bytes of the backward main kernel come from a separate `ncu --set full` capture of this same
    # command (a run under a profiler is never a bench run); the figure is per launch at the weak-scaling
    # shape (B = 8 per GPU) and is omitted for any other shape
    traffic, traffic_source = None, None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists() and Bl == B:
        try:
            tj = json.loads(tf.read_text())
            traffic = tj.get("attn_bwd_main_dram_bytes_per_launch")
            traffic_source = tj.get("source")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "attn_bwd_sm100_persist_kernel<bf16,128>", "achieved": achieved,
                "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained, "traffic": traffic,
                "traffic_source": traffic_source,
                "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)",
                "frac_of_burst": achieved / burst, "frac_of_nominal_2250": achieved / 2250.0,
                "kernel_ms": kern_ms, "algorithmic_flops_per_launch": f_bwd}

    cpu = None
    if not args.no_cpu and world == 1:   # the CPU baseline is timed at N = 1 only (rank 0's host cores)
        threads = os.cpu_count() or 1
        sb, sh, reps = 1, 1, 0
        t_cpu, t_start = 0.0, time.perf_counter()
        cpu_reference_step(sb, sh, threads)   # warm-up (allocator, thread pool)
        while reps < 6 and (time.perf_counter() - t_start) < 20.0:
            t_cpu += cpu_reference_step(sb, sh, threads)
            reps += 1
        cf, cb = flops(sb, sh)
        cpu = {"value": (cf + cb) * reps / t_cpu / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": "port",
               "sample": f"{reps} x (B={sb},H={sh}) units of the C2 shape (L={L},E={E},causal), Float32, "
                         "torch-CPU restatement of the reference's naive attention fwd+bwd"}

    # ---- secondary metric: the HBM-bound ops at config C3's shapes, GB/s vs measured copy BW
    secondary = None
    if not args.no_secondary:
        from nnop_b200 import bwbench
        pk = ROOT / "MEASURED_PEAKS.json"
        hbm = json.loads(pk.read_text()).get("hbm_gbs", 6548.0) if pk.exists() else 6548.0
        secondary = {"unit": "GB/s", "peak": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy)" if pk.exists()
                     else "fallback", "method": "CUDA-graph replay, buffer sets rotated past the 126 MB L2, "
                     "algorithmic bytes of SURVEY.md 8(d)", "ops": bwbench.secondary_block(hbm)}

    line = {
        "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"C2: BF16 causal E=128 L=8192 H=32 B={Bl} fwd+bwd per GPU" +
                               (f" (B=8 in total, batch-sharded over {world})" if strong else ""),
                   "global_batch": Bl * world, "heads": H, "kv_heads": KH, "seq_len": L, "head_dim": E,
                   "parallelism": f"(head,batch)-sharded x{world}, no collective",
                   "l2_policy": f"inputs (4 x {Bl * 67} MB) exceed the 126 MB L2; no flush needed",
                   "flops_convention": "4*B*H*L^2*E/2 fwd, x2.5 bwd (SURVEY.md 8d)"},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": 4 * args.steps * world,
        "gpu_launches_per_step_per_rank": {"attn_fwd_sm100_kernel": 1, "attn_bwd_prep_kernel": 1,
                                           "attn_bwd_sm100_persist_kernel": 1, "attn_bwd_post_kernel": 1},
        "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary,
        "fwd_bwd_tflops_frac_of_sustained_peak": value / world / sustained,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-kv-heads", type=int, default=16, help="kv heads per host-pipeline chunk")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the bandwidth-op block")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: B=8 per GPU (global batch 8*N); strong: the B=8 batch sharded over the N GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
