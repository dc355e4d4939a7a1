"""Interleaved A/B timing of GEMM variants (robust to power/thermal drift): median of several rounds."""
import os
import sys
import statistics

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops

M = 257 * 256
SHAPES = [("proj", M, 1408, 1408, ops.EPI_F32_RESID), ("qkv", M, 4224, 1408, ops.EPI_BF16),
          ("fc1", M, 6144, 1408, ops.EPI_BF16_GELU), ("fc2", M, 1408, 6144, ops.EPI_F32_RESID)]


def timeit(fn, iters=8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    variants = [("ew8", 8), ("ew16", 16)]
    res = {}
    bufs = {}
    for name, m, n, k, epi in SHAPES:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda")
        out = torch.randn(m, n, device="cuda") if epi == ops.EPI_F32_RESID else torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        bufs[name] = (a, w, bias, out)
    for rnd in range(7):
        for name, m, n, k, epi in SHAPES:
            a, w, bias, out = bufs[name]
            for vname, ew in variants:
                ops.gemm_set_epilogue_warps(ew)
                fn = (lambda: ops.gemm(a, w, bias, epi, out=out, resid=out)) if epi == ops.EPI_F32_RESID else \
                     (lambda: ops.gemm(a, w, bias, epi, out=out))
                fn()
                res.setdefault((name, vname), []).append(timeit(fn))
            ops.gemm_set_epilogue_warps(0)
            res.setdefault((name, "cublas"), []).append(timeit(lambda: torch.matmul(a, w.t())))
    for name, m, n, k, epi in SHAPES:
        fl = 2.0 * m * n * k
        line = f"{name:5s}"
        for vname in ("ew8", "ew16", "cublas"):
            ms = statistics.median(res[(name, vname)])
            line += f"  {vname}: {ms:.3f} ms {fl / ms / 1e9:6.0f} TF"
        print(line)


if __name__ == "__main__":
    main()
