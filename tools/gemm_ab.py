"""Interleaved A/B timing of GEMM variants (robust to power/thermal drift): median of several rounds.

    python tools/gemm_ab.py            the six ViT GEMM variants, balanced vs round-robin persistent schedule, next to cuBLAS
                                       (torch.matmul, no epilogue) on the same shapes
    TGB_LIB=tools/_dbg/lib....so python tools/gemm_ab.py     the same with an experiment build of the library
"""
import os
import sys
import statistics

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops

M = 257 * 512
SHAPES = [("proj", M, 1408, 1408, ops.EPI_F32_RESID), ("proj+stats", M, 1408, 1408, ops.EPI_F32_RESID_STATS),
          ("fc2", M, 1408, 6144, ops.EPI_F32_RESID), ("fc2+stats", M, 1408, 6144, ops.EPI_F32_RESID_STATS),
          ("qkv", M, 4224, 1408, ops.EPI_BF16), ("fc1", M, 6144, 1408, ops.EPI_BF16_GELU)]


def timeit(fn, iters=8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    res = {}
    bufs = {}
    for name, m, n, k, epi in SHAPES:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda")
        f32 = epi in (ops.EPI_F32_RESID, ops.EPI_F32_RESID_STATS)
        out = torch.randn(m, n, device="cuda") if f32 else torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        xb = torch.empty(m, n, dtype=torch.bfloat16, device="cuda") if epi == ops.EPI_F32_RESID_STATS else None
        part = torch.empty(ops.ln_stats_slots(m, n), m, 2, device="cuda") if xb is not None else None
        bufs[name] = (a, w, bias, out, xb, part)

    def runner(name, epi):
        a, w, bias, out, xb, part = bufs[name]
        if epi == ops.EPI_F32_RESID:
            return lambda: ops.gemm(a, w, bias, epi, out=out, resid=out)
        if epi == ops.EPI_F32_RESID_STATS:
            return lambda: ops.gemm_ln(a, w, bias, epi, out, resid=out, out_bf16=xb, partial_stats=part)
        return lambda: ops.gemm(a, w, bias, epi, out=out)

    for rnd in range(7):
        for name, m, n, k, epi in SHAPES:
            fn = runner(name, epi)
            for vname, bal in (("groups", 3), ("rowblocks", 2), ("balanced", 1), ("roundrobin", 0)):
                ops.gemm_set_schedule(bal)
                fn()
                res.setdefault((name, vname), []).append(timeit(fn))
            ops.gemm_set_schedule(True)
            a, w = bufs[name][0], bufs[name][1]
            res.setdefault((name, "cublas"), []).append(timeit(lambda: torch.matmul(a, w.t())))
    for name, m, n, k, epi in SHAPES:
        fl = 2.0 * m * n * k
        line = f"{name:10s}"
        for vname in ("groups", "rowblocks", "balanced", "roundrobin", "cublas"):
            if (name, vname) in res:
                ms = statistics.median(res[(name, vname)])
                line += f"  {vname}: {ms:.3f} ms {fl / ms / 1e9:6.0f} TF"
        print(line, flush=True)


if __name__ == "__main__":
    main()
