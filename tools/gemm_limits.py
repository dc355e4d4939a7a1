"""What bounds the tcgen05 GEMM under sustained load?  Times the four ViT GEMM shapes with experiment builds of the
kernel (tools/_dbg/libtgb_x{1,2,3}.so, built with -DTGB_GEMM_X=n: bit 0 = no TMA loads after the ring is primed,
bit 1 = empty epilogue) next to the product library; ~1.5 s of back-to-back launches per point, SM clock sampled while
the GPU is busy.  Results of the experiment builds are numerically meaningless by construction.

    python tools/gemm_limits.py            (spawns one subprocess per library)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(lib):
    import torch
    import pynvml
    from videotgb_b200 import _lib
    if lib != "product":
        _lib.LIB_PATH = os.path.join(ROOT, "tools", "_dbg", lib)
    from videotgb_b200 import ops
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    m = 257 * 512
    shapes = [("qkv", m, 4224, 1408, ops.EPI_BF16), ("fc1", m, 6144, 1408, ops.EPI_BF16_GELU),
              ("fc2", m, 1408, 6144, ops.EPI_F32_RESID), ("proj", m, 1408, 1408, ops.EPI_F32_RESID)]
    out = {}
    for name, m, n, k, epi in shapes:
        a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
        w = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda")
        o = torch.randn(m, n, device="cuda") if epi == ops.EPI_F32_RESID else torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
        fn = (lambda: ops.gemm(a, w, bias, epi, out=o, resid=o)) if epi == ops.EPI_F32_RESID else (lambda: ops.gemm(a, w, bias, epi, out=o))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        iters = max(10, int(1500.0 / e0.elapsed_time(e1)))
        for _ in range(iters // 2):                      # bring the chip to its sustained power state
            fn()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name] = {"ms": round(ms, 4), "tflops": round(2.0 * m * n * k / ms / 1e9, 1), "sm_mhz": clk, "power_w": round(pw)}
        del a, w, o
    print(json.dumps({"lib": lib, "res": out}), flush=True)


def main():
    if len(sys.argv) > 1:
        return child(sys.argv[1])
    for lib in ("product", "libtgb_x1.so", "libtgb_x2.so", "libtgb_x3.so"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), lib], capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)


if __name__ == "__main__":
    main()
