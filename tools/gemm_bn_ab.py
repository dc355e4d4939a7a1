"""Interleaved A/B of the N-tile width (bn = 256 / 128) for the two residual GEMMs of a ViT layer."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops
from tools.gemm_ab import timeit

M = 257 * 512
SHAPES = [("proj", M, 1408, 1408, ops.EPI_F32_RESID), ("fc2", M, 1408, 6144, ops.EPI_F32_RESID),
          ("fc1", M, 6144, 1408, ops.EPI_BF16_GELU)]


def main():
    res, bufs = {}, {}
    for name, m, n, k, epi in SHAPES:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        out = torch.randn(m, n, device="cuda") if epi == ops.EPI_F32_RESID else torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        bufs[name] = (a, w, torch.randn(n, device="cuda"), out)
    for rnd in range(7):
        for name, m, n, k, epi in SHAPES:
            a, w, bias, out = bufs[name]
            for bn in (0, 128):
                fn = (lambda: ops.gemm(a, w, bias, epi, out=out, resid=out, bn=bn)) if epi == ops.EPI_F32_RESID else \
                     (lambda: ops.gemm(a, w, bias, epi, out=out, bn=bn))
                try:
                    fn()
                except Exception as e:                      # noqa: BLE001
                    res[(name, bn)] = [float("nan")]
                    print(name, bn, "unsupported:", str(e)[:80])
                    continue
                res.setdefault((name, bn), []).append(timeit(fn))
    for name, m, n, k, epi in SHAPES:
        fl = 2.0 * m * n * k
        print(name, "  ".join(f"bn={bn or 256}: {statistics.median(res[(name, bn)]):.3f} ms {fl / statistics.median(res[(name, bn)]) / 1e9:6.0f} TF"
                              for bn in (0, 128)))


if __name__ == "__main__":
    main()
