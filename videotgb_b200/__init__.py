"""videotgb_b200 -- B200-native (sm_100a) implementation of the VideoTGB data-parallel hot path.

    temporal sampler scoring -> Gumbel span draw / frame selection / gather -> EVA-ViT-g -> BLIP-2 / InstructBLIP
    Q-Former -> language_projection -> LLM query tokens

Hand-written CUDA (tcgen05 + TMA GEMMs, fused attention, HBM-bound frame kernels) behind a C ABI
(include/videotgb_b200.h, libvideotgb_b200.so), with a Python host layer that mirrors the reference module API.
There is no CPU / PyTorch fallback: without the built library or off a B200 every op raises.
"""
__version__ = "0.1.0"

__all__ = ["ops", "modeling", "sampler", "selection", "transforms", "pipeline", "synthetic", "build"]
