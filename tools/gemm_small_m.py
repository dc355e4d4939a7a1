"""Tile choice for small-M GEMMs (a 4..32-frame ViT pass): time every (N tile, cta_group) the library supports on the
qkv / fc1 shapes, as CUDA-graph replays (no host time between launches), next to the library's own choice.

    python tools/gemm_small_m.py        (B200 box)
"""
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
    return e0.elapsed_time(e1) / (iters * replays) * 1e3      # us


def main():
    for frames in (4, 8, 16, 32, 64):
        m = 257 * frames
        shapes = (("qkv", 4224, 1408, ops.EPI_BF16), ("fc1", 6144, 1408, ops.EPI_BF16_GELU))
        if os.environ.get("TGB_SMALL_M_RESID"):      # the residual GEMMs (the statistics variants share their tiles and schedule)
            shapes = (("proj", 1408, 1408, ops.EPI_F32_RESID), ("fc2", 1408, 6144, ops.EPI_F32_RESID))
        for name, n, k, epi in shapes:
            a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
            # several weight copies, cycled, so that every launch streams its weights from DRAM like a real layer stack does
            ws = [(torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16) for _ in range(10)]
            bias = torch.randn(n, device="cuda")
            resid = epi == ops.EPI_F32_RESID
            out = torch.randn(m, n, device="cuda") if resid else torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
            line = f"M={m:6d} {name}:"
            state = {"i": 0}

            def run(bn):
                w = ws[state["i"] % len(ws)]
                state["i"] += 1
                if resid:
                    ops.gemm(a, w, bias, epi, out=out, resid=out, bn=bn)
                else:
                    ops.gemm(a, w, bias, epi, out=out, bn=bn)

            for cg in (2, 1):
                ops.gemm_set_cta_group(cg)
                for bn in (256, 128):
                    line += f"  cg{cg}/bn{bn} {graph_time(lambda: run(bn)):6.1f} us"
            ops.gemm_set_cta_group(0)
            line += f"  | auto {graph_time(lambda: run(0)):6.1f} us"
            print(line, flush=True)


if __name__ == "__main__":
    main()
