"""ln_stats_finalize on a 512-frame pass (M = 131584 rows): time per launch by slot count (CUDA-graph replays)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops


def graph_time(fn, iters=20, replays=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * replays)


m, n = 257 * 512, 1408
stats = torch.empty(m, 2, device="cuda")
ratio = torch.zeros(1, device="cuda")
for slots in (12, 22):
    parts = [torch.randn(slots, m, 2, device="cuda") for _ in range(8)]      # rotate: not L2 resident
    i = [0]

    def run():
        ops.ln_stats_finalize(parts[i[0] % 8], n, 1e-6, stats, ratio)
        i[0] += 1

    ms = graph_time(run, iters=16)
    print(f"slots {slots}: {ms * 1e3:6.2f} us per launch, {slots * m * 8 / ms / 1e6:6.0f} GB/s")
