import os, sys
sys.path.insert(0, os.getcwd())
import torch
from videotgb_b200 import ops
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (t, h, w) in [(512, 360, 640), (256, 480, 854), (64, 1080, 1920), (512, 240, 320)]:
    native = torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, device="cuda")
    for var in (0, 1):
        ops.resize_set_variant(var)
        ms = timeit(lambda: ops.resize_bilinear_u8(native, 224))
        nb = t * (h * w * 3 + 150528)
        print(f"{h}x{w} variant {var}: {ms:.3f} ms  {nb / ms / 1e6:.0f} GB/s", flush=True)
    ops.resize_set_variant(0)
