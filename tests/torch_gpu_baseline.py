"""Context number: the reference's own PyTorch path on the SAME B200 -- the oracle port run on the
GPU under torch.autocast(bf16), i.e. Lightning `precision: bf16` as the reference trains / evaluates -- next to this
repo's kernels on the same frames.  Not a pytest module (run by hand on the GPU box):

    python tests/torch_gpu_baseline.py [frames=256]  > gpurun_out/torch_gpu_baseline.json

Lives under tests/ because it executes oracle/ (test infrastructure); nothing in the product imports it.
"""
import json
import os
import sys
from types import SimpleNamespace as NS

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import torch_ref
from videotgb_b200 import synthetic as syn
from videotgb_b200.modeling import Blip2QFormerModel, Blip2VisionModel


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
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = "cuda"
    vsd = {k: v.to(dev) for k, v in syn.vision_state_dict().items()}
    qsd = {k: v.to(dev) for k, v in syn.qformer_state_dict().items()}
    hsd = {k: v.to(dev) for k, v in syn.head_state_dict().items()}
    px = torch.randn(frames, 3, 224, 224, device=dev)
    qt = hsd["query_tokens"].expand(frames, -1, -1)

    def torch_path():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            img, _ = torch_ref.vit_forward(vsd, px)
            return torch_ref.qformer_forward(qsd, qt, img, None)

    vit = Blip2VisionModel(NS()).to(dev)
    vit.load_state_dict(vsd)
    qf = Blip2QFormerModel(None).to(dev)
    qf.load_state_dict(qsd)

    def ours():
        with torch.no_grad():
            img = vit(pixel_values=px, return_dict=True).last_hidden_state
            return qf(query_embeds=qt, encoder_hidden_states=img)[0]

    a, b = torch_path().float(), ours().float()
    rel = ((a - b).abs().max() / b.abs().max()).item()
    ms_t, ms_o = timed(torch_path), timed(ours)
    print(json.dumps({"what": "ViT-g (39 layers) + BLIP-2 Q-Former on one B200, same frames and weights",
                      "frames": frames, "torch_autocast_bf16_ms": ms_t, "torch_frames_per_s": frames / ms_t * 1e3,
                      "videotgb_b200_ms": ms_o, "videotgb_b200_frames_per_s": frames / ms_o * 1e3,
                      "speedup": ms_t / ms_o, "max_rel_diff_between_the_two": rel}))


if __name__ == "__main__":
    main()
