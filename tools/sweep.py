"""BASELINE config 5: throughput sweep, frames/video x batch, through the whole hot path (sampler -> select -> gather ->
ViT-g -> Q-Former -> projection) on one GPU.  Writes one JSON line per point.  Run on the B200 box:

    python tools/sweep.py > gpurun_out/r1_sweep.jsonl
"""
import json
import os
import sys
from types import SimpleNamespace as NS

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200.pipeline import VideoTGBEncoder

POINTS = [(8, 1), (8, 4), (8, 16), (8, 64), (32, 1), (32, 4), (32, 16), (32, 64), (128, 1), (128, 4), (128, 16),
          (512, 1), (512, 4), (2048, 1)]


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    with torch.device(dev):
        enc = VideoTGBEncoder(NS(), None, NS(fusion_layer=6, encoder_width=768, max_position_embeddings=2112), 2048)
        with torch.no_grad():
            enc.model.query_tokens.normal_(0.0, 0.02)
    enc.eval()
    gen = torch.Generator(device=dev).manual_seed(5)
    for k, b in POINTS:
        n = k                                   # candidates == flow frames == frames/video; the sampler picks k of them
        lof = k
        batch = {
            "frames_u8": torch.randint(0, 256, (b, n, 224, 224, 3), dtype=torch.uint8, device=dev),
            "of": torch.randn(b, lof, 2, 224, 224, device=dev) * 0.25,
            "of_mask": torch.ones(b, lof + 2, dtype=torch.long, device=dev),
            "of_lengths": torch.full((b,), lof, dtype=torch.int32, device=dev),
            "sampler_question": torch.randint(1000, 30522, (b, 16), device=dev),
            "sampler_question_attention_mask": torch.ones(b, 16, dtype=torch.long, device=dev),
            "nframe": k,
        }
        for _ in range(2):
            enc.encode_u8(batch, generator=gen)
        torch.cuda.synchronize()
        steps = 3 if b * k >= 512 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            enc.encode_u8(batch, generator=gen)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"frames_per_video": k, "videos": b, "frames": b * k, "ms_per_step": round(ms, 3),
                          "frames_per_s": round(b * k / ms * 1e3, 1)}), flush=True)
        del batch
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
