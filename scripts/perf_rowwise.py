"""HBM-bound ops vs the measured copy bandwidth (MEASURED_PEAKS.json): graph-replayed, L2-rotated timing
(nnop_b200/bwbench.py) at BASELINE config C3's shapes, the reference's benchmark shapes
(benchmarks/main.jl:70-300) and large streaming shapes.   python scripts/perf_rowwise.py [--quick]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nnop.jl_b200"))
import torch  # noqa: E402
from nnop_b200 import bwbench as BW  # noqa: E402

P = ROOT / "MEASURED_PEAKS.json"
PEAK = json.loads(P.read_text())["hbm_gbs"] if P.exists() else 6548.0


def show(rows):
    for r in rows:
        print(f"{r['op']:15s} {r['dtype']:9s} {r['shape']:34s} {r['us']:9.2f} us {r['gbs']:8.1f} GB/s "
              f"{100 * r['gbs'] / PEAK:5.1f}% of measured copy BW", flush=True)
    torch.cuda.empty_cache()


def main():
    quick = "--quick" in sys.argv
    print(f"# graph-replayed, buffer sets rotated past L2; peak = {PEAK:.0f} GB/s (measured copy)")
    for dtype in (torch.bfloat16, torch.float32):
        for n, emb in ((8192, 4096), (65536, 4096), (1024, 1024)) if not quick else ((8192, 4096),):
            show(BW.norm_ops(dtype, n, emb))
        for cols, N in ((1024, 8192), (16384, 8192)) if not quick else ((1024, 8192),):
            show(BW.softmax_ops(dtype, cols, N))
        show(BW.rope_op(dtype, 1, 32, 8, 8192, 128))
        if not quick:
            show(BW.rope_op(dtype, 8, 32, 8, 8192, 128))
            show(BW.rope_op(dtype, 4, 3, 3, 1024, 64))   # benchmarks/main.jl:189-261


if __name__ == "__main__":
    main()
