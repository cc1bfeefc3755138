"""Copy-only diagnosis of the host-buffer (e2e) path: H2D alone, D2H alone and both at once, per rank and
summed over ranks, with bench.py's e2e volumes (2 GiB each way per step at config C2) and no kernels.
Separates host-memory / PCIe limits from anything the kernels do (VERDICT r01 "Next round" 5).

    python scripts/perf_copy_only.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/perf_copy_only.py     # N ranks copying at the same time
"""
import os
import sys

import torch


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 2 << 30
    results = {}
    for label, chunk in (("one 2 GiB copy", nbytes), ("64 x 32 MiB copies", 32 << 20)):
        host_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host_in.fill_(1)
        host_out.fill_(0)
        dev_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        dev_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def h2d():
            with torch.cuda.stream(s1):
                for o in range(0, nbytes, chunk):
                    dev_in[o:o + chunk].copy_(host_in[o:o + chunk], non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                for o in range(0, nbytes, chunk):
                    host_out[o:o + chunk].copy_(dev_out[o:o + chunk], non_blocking=True)

        for name, fns in (("H2D", (h2d,)), ("D2H", (d2h,)), ("duplex", (h2d, d2h))):
            for f in fns:
                f()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            s1.wait_event(a)
            s2.wait_event(a)
            reps = 3
            for _ in range(reps):
                for f in fns:
                    f()
            cur = torch.cuda.current_stream()
            cur.wait_stream(s1)
            cur.wait_stream(s2)
            b.record()
            torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
            if dist is not None:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            results[(label, name)] = nbytes / ms.item() / 1e6   # GB/s per direction per rank (slowest rank)
        del host_in, host_out, dev_in, dev_out
    if rank == 0:
        print(f"# copy-only, {world} rank(s) at once, pinned host memory, 2 GiB per direction per rank; GB/s per direction")
        for (label, name), gbs in results.items():
            print(f"{label:20s} {name:7s} {gbs:7.1f} GB/s per rank   {gbs * world:8.1f} GB/s all ranks", flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
