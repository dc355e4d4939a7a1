"""Interleaved A/B of the fp32-residual epilogue's piece width (16 vs 32 columns) on the four residual GEMMs of a ViT pass.

    python tools/gemm_piece_ab.py      (B200 box; medians of 7 interleaved rounds, CUDA events)
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops

M = 257 * 512
SHAPES = [("proj", M, 1408, 1408, ops.EPI_F32_RESID), ("proj+stats", M, 1408, 1408, ops.EPI_F32_RESID_STATS),
          ("fc2", M, 1408, 6144, ops.EPI_F32_RESID), ("fc2+stats", M, 1408, 6144, ops.EPI_F32_RESID_STATS)]


def timeit(fn, iters=8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res, bufs = {}, {}
    for name, m, n, k, epi in SHAPES:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda")
        out = torch.randn(m, n, device="cuda")
        xb = torch.empty(m, n, dtype=torch.bfloat16, device="cuda") if epi == ops.EPI_F32_RESID_STATS else None
        part = torch.empty(ops.ln_stats_slots(m, n), m, 2, device="cuda") if xb is not None else None
        bufs[name] = (a, w, bias, out, xb, part)

    def runner(name, epi):
        a, w, bias, out, xb, part = bufs[name]
        if epi == ops.EPI_F32_RESID:
            return lambda: ops.gemm(a, w, bias, epi, out=out, resid=out)
        return lambda: ops.gemm_ln(a, w, bias, epi, out, resid=out, out_bf16=xb, partial_stats=part)

    for rnd in range(7):
        for name, m, n, k, epi in SHAPES:
            fn = runner(name, epi)
            for pw in (16, 32):
                ops.gemm_set_resid_piece(pw)
                fn()
                res.setdefault((name, pw), []).append(timeit(fn))
            ops.gemm_set_resid_piece(0)
    for name, m, n, k, epi in SHAPES:
        fl = 2.0 * m * n * k
        line = f"{name:10s}"
        for pw in (16, 32):
            ms = statistics.median(res[(name, pw)])
            line += f"  pieces of {pw} columns: {ms:.3f} ms {fl / ms / 1e9:6.0f} TF/s"
        print(line, flush=True)


if __name__ == "__main__":
    main()
