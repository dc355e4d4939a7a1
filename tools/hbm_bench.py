"""Achieved HBM bandwidth of the memory-bound kernels (north_star: >= 70 % of the measured copy bandwidth).

    python tools/hbm_bench.py            (B200 box; prints one JSON line per kernel, algorithmic bytes / CUDA-event time)

Buffers are sized well above the 126 MB L2 so every iteration streams from HBM.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import ops

PEAK = 6550.1
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        PEAK = json.load(f)["hbm_gbs"]
except OSError:
    pass


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, nbytes, ms, unit):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "algorithmic_bytes": nbytes, "ms": round(ms, 4), "GB/s": round(gbs, 1),
                      "frac_of_measured_peak": round(gbs / PEAK, 3), "peak_GB/s": PEAK, "unit_of_work": unit}), flush=True)


def main():
    dev = "cuda"
    f = 2048
    u8 = torch.randint(0, 256, (f, 224, 224, 3), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: ops.normalize_frames_u8(u8))
    report("normalize_u8_kernel (u8 THWC -> f32 TCHW)", f * (150528 + 602112), ms, f"{f} frames")
    ms = timeit(lambda: ops.patchify_u8(u8, 14, 592))
    report("patchify_u8_pair_kernel (normalise + im2col -> bf16)", f * (150528 + 256 * 592 * 2), ms, f"{f} frames")
    ops.patchify_set_variant(1)
    ms = timeit(lambda: ops.patchify_u8(u8, 14, 592))
    ops.patchify_set_variant(0)
    report("patchify_kernel<u8> (generic fallback: residual-corrected normalise)", f * (150528 + 256 * 592 * 2), ms, f"{f} frames")
    native = torch.randint(0, 256, (512, 360, 640, 3), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: ops.resize_bilinear_u8(native, 224))
    report("resize_bilinear_u8_kernel (360x640 -> 224x224, u8)", 512 * (360 * 640 * 3 + 150528), ms, "512 decoded frames")
    del native
    px = torch.randn(1024, 3, 224, 224, device=dev)
    ms = timeit(lambda: ops.patchify(px, 14, 592))
    report("patchify_kernel<f32> (im2col -> bf16)", 1024 * (602112 + 256 * 592 * 2), ms, "1024 frames")
    b, n, k = 32, 64, 32
    src = torch.randn(b, n, 3, 224, 224, device=dev)
    idx = torch.sort(torch.randint(0, n, (b, k), device=dev), dim=1).values
    ms = timeit(lambda: ops.gather_frames(src, idx))
    report("gather_frames_kernel (fp32 frames, reference dtype)", b * k * 2 * 602112, ms, f"{b * k} selected frames")
    src8 = u8.view(32, 64, 224, 224, 3)
    ms = timeit(lambda: ops.gather_frames(src8, idx))
    report("gather_frames_kernel (u8 frames)", b * k * 2 * 150528, ms, f"{b * k} selected frames")
    flow = torch.randn(1024, 2, 224, 224, device=dev)
    fcw = torch.randn(196, device=dev)
    ms = timeit(lambda: ops.of_patch_reduce(flow, fcw))
    report("of_patch_reduce_kernel (flow embedding)", 1024 * 401408, ms, "1024 candidate flow frames")
    x = torch.randn(65792, 1408, device=dev)
    g, bb = torch.ones(1408, device=dev), torch.zeros(1408, device=dev)
    out = torch.empty(65792, 1408, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.layernorm(x, g, bb, 1e-6, out_bf16=out))
    report("layernorm_kernel (f32 -> bf16, N=1408)", 65792 * 1408 * 6, ms, "65792 rows")
    ms = timeit(lambda: ops.cast_bf16(x))
    report("cast_f32_bf16_kernel", 65792 * 1408 * 6, ms, "65792 rows")


if __name__ == "__main__":
    main()
