"""Where the time of a SMALL ViT pass goes (BASELINE config 4 on 8 GPUs leaves 4 frames per rank): times the 39-layer
EVA-ViT-g forward on f frames eagerly and as a CUDA-graph replay; under `ncu --metrics gpu__time_duration.sum` the eager
pass gives the per-kernel list.

    python tools/small_batch_profile.py [frames ...]          (B200 box)
"""
import json
import os
import sys
from types import SimpleNamespace as NS

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200.modeling import Blip2VisionModel


def timeit(fn, iters=10):
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


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    with torch.device(dev):
        m = Blip2VisionModel(NS(num_hidden_layers=39))
    m.eval()
    for f in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]:
        u8 = torch.randint(0, 256, (f, 224, 224, 3), dtype=torch.uint8, device=dev)
        m.graph_max_rows = 0
        eager = timeit(lambda: m.forward_u8(u8))
        m.graph_max_rows = 64
        graph = timeit(lambda: m.forward_u8(u8))
        print(json.dumps({"frames": f, "eager_ms": round(eager, 3), "graph_ms": round(graph, 3),
                          "weights_hbm_floor_ms": round(2.0e9 / 6.55e12 * 1e3, 3)}), flush=True)


if __name__ == "__main__":
    main()
