"""First-contact probe + micro-benchmark for the tcgen05 GEMM (run on the B200 box).

    python tools/gemm_probe.py [--bench]

Prints, per shape, the max error vs an fp64 reference and -- on mismatch -- where the error lives (which rows /
columns / K-slices), which is what one needs to debug descriptor or pipeline mistakes without a local GPU.
"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops


def probe(m, n, k, bn, epi=ops.EPI_F32):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    w = (torch.randn(n, k, generator=g) * 0.1).to(torch.bfloat16)
    want = a.double() @ w.double().t()
    try:
        got = ops.gemm(a.cuda(), w.cuda(), None, epi, bn=bn)
        torch.cuda.synchronize()
    except Exception as e:                                  # noqa: BLE001
        print(f"[{m}x{n}x{k} bn={bn}] FAILED: {e}")
        return False
    got = got.cpu().double()
    err = (got - want).abs()
    tol = 1e-3 * want.abs().max().item()
    ok = err.max().item() <= tol
    print(f"[{m}x{n}x{k} bn={bn}] max|err|={err.max().item():.3e} (tol {tol:.3e}) {'OK' if ok else 'MISMATCH'}")
    if not ok:
        bad = err > tol
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows:", rows[:16].tolist(), "... n=", rows.numel(), " bad cols:", cols[:16].tolist(), "... n=", cols.numel())
        print("   got[0,:8] ", got[0, :8].tolist())
        print("   want[0,:8]", want[0, :8].tolist())
        # which single k-slice of 16 / 64 explains the result?  (helps spot descriptor-advance bugs)
        for width in (16, 64):
            for ks in range(0, min(k, 256), width):
                part = a[:, ks:ks + width].double() @ w[:, ks:ks + width].double().t()
                d = (got - part).abs().max().item()
                if d < tol:
                    print(f"   output equals the K[{ks}:{ks + width}] partial product only")
        if got.abs().max().item() == 0:
            print("   output is all zeros")
    return ok


def bench(m, n, k, epi, bn=0, iters=20):
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    resid = torch.randn(m, n, device="cuda") if epi == ops.EPI_F32_RESID else None
    out = resid if resid is not None else None
    for _ in range(3):
        ops.gemm(a, w, bias, epi, out=out, resid=resid, bn=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, bias, epi, out=out, resid=resid, bn=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * m * n * k / ms / 1e9
    # cuBLAS reference point (same shapes, torch.matmul)
    for _ in range(3):
        torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"bench M={m} N={n} K={k} epi={epi} bn={bn}: {ms:.3f} ms  {tf:.0f} TFLOP/s   (cuBLAS matmul {2.0 * m * n * k / ms2 / 1e9:.0f} TFLOP/s)")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    ok = True
    for cg in (2, 1):
        ops.gemm_set_cta_group(cg)
        print("cta_group", cg)
        for shape in [(256, 256, 64, 256), (256, 512, 256, 256), (300, 264, 200, 256), (520, 4224, 1408, 128)]:
            ok = probe(*shape) and ok
    ops.gemm_set_cta_group(0)
    for shape in [(128, 64, 64, 64), (128, 256, 64, 256), (128, 256, 256, 256), (256, 512, 512, 256), (300, 264, 200, 256),
                  (1028, 1408, 1408, 128)]:
        ok = probe(*shape) and ok
    if "--bench" in sys.argv and ok:
        M = 257 * 256
        for ew in (8, 16):
            ops.gemm_set_epilogue_warps(ew)
            print("cta_group 2, epilogue warps", ew)
            bench(M, 1408, 1408, ops.EPI_F32_RESID, 256)
            bench(M, 4224, 1408, ops.EPI_BF16, 256)
            bench(M, 6144, 1408, ops.EPI_BF16_GELU, 256)
            bench(M, 1408, 6144, ops.EPI_F32_RESID, 256)
            bench(8192, 8192, 8192, ops.EPI_BF16, 256)
        ops.gemm_set_epilogue_warps(0)
