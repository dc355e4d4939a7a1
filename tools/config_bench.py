"""BASELINE configs 3 and 4 at full model size on one GPU (random-init weights, synthetic inputs; the parity of these
paths is covered by tests/test_modules_gpu.py and tests/test_sharded_gpu.py -- this only times them).

  config 3  InstructBLIP-Vicuna-7B path: one 128-frame long video, the sampler (mode "multi_modal", 12 layers) scores
            all 128 flow frames, 32 frames are selected ("ratio" mapping), instruction-conditioned Q-Former,
            mean-pooled tokens [1, 32, 4096]; optionally with the RAFT flow extraction over the 127 frame pairs in front
  config 4  ActivityNet-QA-shaped long video: 1024 candidate frames scored, span selection, 32 frames gathered and
            encoded (the single-GPU form of `encode_sharded_u8`; the 8-GPU form shards the 1024 candidates by frame)

    python tools/config_bench.py > gpurun_out/r1_configs_3_4.jsonl
"""
import json
import os
import sys
from types import SimpleNamespace as NS

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200.pipeline import VideoTGBEncoder


def timeit(fn, iters):
    for _ in range(2):
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
    torch.manual_seed(3)
    gen = torch.Generator(device=dev).manual_seed(9)
    # ---------------------------------------------------------------- config 3
    t, k, tq, tt = 128, 32, 16, 24
    with torch.device(dev):
        enc = VideoTGBEncoder(NS(), None, NS(fusion_layer=6, encoder_width=768, max_position_embeddings=2112), 4096,
                              instruct=True, with_flow_extractor=True)
        with torch.no_grad():
            enc.model.query_tokens.normal_(0.0, 0.02)
    enc.eval()
    frames = torch.randint(0, 256, (1, t, 224, 224, 3), dtype=torch.uint8, device=dev)
    batch = {
        "frames_u8": frames, "of": torch.randn(1, t, 2, 224, 224, device=dev) * 0.25,
        "of_mask": torch.ones(1, t + 2, dtype=torch.long, device=dev), "of_lengths": [t],
        "sampler_question": torch.randint(1000, 30522, (1, tq), device=dev),
        "sampler_question_attention_mask": torch.ones(1, tq, dtype=torch.long, device=dev),
        "qformer_text": torch.randint(1000, 30522, (1, tt), device=dev),
        "qformer_text_attention_mask": torch.ones(1, tt, dtype=torch.long, device=dev), "nframe": k,
    }
    ms = timeit(lambda: enc.encode_u8(batch, mode="multi_modal", variant="ratio", assemble="mean", generator=gen), 10)
    print(json.dumps({"config": 3, "what": "InstructBLIP path, 128-frame video -> 32 frames, flow given", "ms": round(ms, 3),
                      "selected_frames_per_s": round(k / ms * 1e3, 1), "candidate_frames_per_s": round(t / ms * 1e3, 1)}), flush=True)
    rgb = frames[0].permute(0, 3, 1, 2).float().contiguous()          # [T,3,H,W], 0..255 as the reference feeds RAFT

    def with_flow():
        b2 = dict(batch)
        b2["of"] = enc.extract_flow([rgb])
        return enc.encode_u8(b2, mode="multi_modal", variant="ratio", assemble="mean", generator=gen)

    ms2 = timeit(with_flow, 5)
    print(json.dumps({"config": 3, "what": "the same with RAFT flow extraction over the 127 frame pairs in front", "ms": round(ms2, 3),
                      "raft_ms": round(ms2 - ms, 3), "candidate_frames_per_s": round(t / ms2 * 1e3, 1)}), flush=True)
    del enc, batch, frames, rgb
    torch.cuda.empty_cache()
    # ---------------------------------------------------------------- config 4 (single-GPU form)
    t, k = 1024, 32
    with torch.device(dev):
        enc = VideoTGBEncoder(NS(), None, NS(fusion_layer=6, encoder_width=768, max_position_embeddings=2112), 2048)
        with torch.no_grad():
            enc.model.query_tokens.normal_(0.0, 0.02)
    enc.eval()
    batch = {
        "frames_u8": torch.randint(0, 256, (1, t, 224, 224, 3), dtype=torch.uint8, device=dev),
        "of": torch.randn(1, t, 2, 224, 224, device=dev) * 0.25,
        "of_mask": torch.ones(1, t + 2, dtype=torch.long, device=dev), "of_lengths": [t],
        "sampler_question": torch.randint(1000, 30522, (1, 16), device=dev),
        "sampler_question_attention_mask": torch.ones(1, 16, dtype=torch.long, device=dev), "nframe": k,
    }
    ms = timeit(lambda: enc.encode_u8(batch, generator=gen), 10)
    ms_score = timeit(lambda: enc.score(batch, "fusion"), 10)
    print(json.dumps({"config": 4, "what": "1024 candidate frames scored (fusion sampler), 32 selected and encoded, one GPU",
                      "ms": round(ms, 3), "score_ms": round(ms_score, 3), "candidate_frames_per_s": round(t / ms * 1e3, 1),
                      "candidates_scored_per_s": round(t / ms_score * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
