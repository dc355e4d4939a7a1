"""Context numbers for the RAFT extractor: this repo's kernels vs the reference algorithm in
plain torch (oracle port, fp32 and TF32-free) on the same B200, 31 frame pairs of 224x224 (a 32-frame video).
Run by hand on the GPU box:  python tests/raft_gpu_baseline.py [pairs] > gpurun_out/raft_baseline.json
Lives under tests/ because it executes oracle/ (test infrastructure)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import raft_ref
from tests.parity_utils import cosine, relerr
from videotgb_b200 import ops
from videotgb_b200 import synthetic as syn
from videotgb_b200.raft import RAFT


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 31
    dev = "cuda"
    g = torch.Generator().manual_seed(5)
    base = torch.nn.functional.interpolate(torch.rand(1, 3, 30, 30, generator=g), size=(224 + 2 * pairs, 224 + 2 * pairs),
                                           mode="bicubic", align_corners=False).clamp(0, 1) * 255.0
    frames = torch.cat([base[:, :, i: i + 224, i: i + 224] for i in range(pairs + 1)], 0).contiguous().to(dev)
    sd = {k: v.to(dev) for k, v in syn.raft_state_dict().items()}
    m = RAFT().to(dev)
    m.load_state_dict(sd)
    m.eval()

    def torch_path():
        with torch.no_grad():
            return raft_ref.raft_forward(sd, frames[:-1], frames[1:], iters=20)

    def ours():
        return m(frames[:-1], frames[1:])

    a, b = torch_path(), ours()
    l0 = ops.LAUNCHES
    ours()
    launches = ops.LAUNCHES - l0
    ms_t, ms_o = timed(torch_path), timed(ours)
    print(json.dumps({"what": "RAFT (20 iterations), 224x224 frame pairs on one B200, same weights", "pairs": pairs,
                      "torch_fp32_ms": ms_t, "videotgb_b200_ms": ms_o, "speedup": ms_t / ms_o,
                      "pairs_per_s": pairs / ms_o * 1e3, "kernel_launches": launches,
                      "relerr_vs_torch_fp32": relerr(b, a), "cosine": cosine(b, a), "flow_absmax": float(a.abs().max())}))


if __name__ == "__main__":
    main()
