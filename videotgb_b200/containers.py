"""The container objects the reference task modules build and call (`self.model`), backed by the sm_100a kernels.

    Blip2ForConditionalGeneration            <- src/models/components/xblip2.py:1539-1556
    InstructBlipForConditionalGeneration     <- src/models/components/xinstructblip.py:1255-1284

Both own `vision_model`, `query_tokens`, `qformer`, `language_projection`, `temporal_projection` (present in the
reference's checkpoints, every use is commented out there), `config`, `language_model` and `get_input_embeddings()`,
which is everything `eval_forward` touches (LSTP_SF_blip2_module.py:405-506, LSTP_SF_module.py:489-580).  Parameter
names and shapes are the reference's, so `load_state_dict` / `from_pretrained` of a reference (= HF BLIP-2 /
InstructBLIP) checkpoint fills them.

The language model is NOT part of the hot path: it is the unchanged HF model.  It is built from `config.text_config`
exactly as the reference does (`AutoModelForCausalLM` / `AutoModelForSeq2SeqLM.from_config`), or injected by the caller
(`language_model=`), or left out (`language_model=False`) when only the visual tokens are wanted.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict, Iterable, Optional

import torch
from torch import nn

from .modeling import (Blip2QFormerModel, Blip2VisionModel, InstructBlipQFormerModel, InstructBlipVisionModel,
                       _KernelModule, _Linear, _cfg)
from . import ops


# ------------------------------------------------------------------------------------------------ checkpoints
def read_checkpoint(path: str, prefixes: Optional[Iterable[str]] = None) -> Dict[str, torch.Tensor]:
    """State dict of an HF-style checkpoint DIRECTORY (`model.safetensors`, a sharded `*.index.json`, or
    `pytorch_model.bin`) or of a single `.safetensors` / `.bin` / `.pt` / `.ckpt` file (Lightning checkpoints:
    the `state_dict` entry, eval/utils/builder_utils.py:178-180).  `prefixes`: keep only keys starting with one of
    them (the LLM shards of a BLIP-2 checkpoint are then never read)."""
    keep = tuple(prefixes) if prefixes is not None else None

    def want(k):
        return keep is None or k.startswith(keep)

    def load_file(fp):
        if fp.endswith(".safetensors"):
            from safetensors import safe_open
            out = {}
            with safe_open(fp, framework="pt", device="cpu") as f:
                for k in f.keys():
                    if want(k):
                        out[k] = f.get_tensor(k)
            return out
        sd = torch.load(fp, map_location="cpu", weights_only=False)
        if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
            sd = sd["state_dict"]
        return {k: v for k, v in sd.items() if want(k)}

    if os.path.isfile(path):
        return load_file(path)
    if not os.path.isdir(path):
        try:                                    # a hub id: resolve it to a local snapshot (needs network or a warm cache)
            from huggingface_hub import snapshot_download
            path = snapshot_download(path)
        except Exception as e:                  # noqa: BLE001
            raise OSError(f"{path!r} is neither a checkpoint file, a directory, nor a resolvable hub id: {e}") from e
    for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
        ip = os.path.join(path, index)
        if os.path.exists(ip):
            with open(ip) as f:
                weight_map = json.load(f)["weight_map"]
            out = {}
            for shard in sorted({s for k, s in weight_map.items() if want(k)}):
                out.update(load_file(os.path.join(path, shard)))
            return out
    for name in ("model.safetensors", "pytorch_model.bin"):
        fp = os.path.join(path, name)
        if os.path.exists(fp):
            return load_file(fp)
    raise OSError(f"no model.safetensors / pytorch_model.bin (or sharded index) under {path}")


def _report(missing, unexpected, what):
    if missing or unexpected:
        import warnings
        warnings.warn(f"{what}: missing keys {sorted(missing)[:8]}{'...' if len(missing) > 8 else ''}; "
                      f"unexpected keys {sorted(unexpected)[:8]}{'...' if len(unexpected) > 8 else ''}")


# ------------------------------------------------------------------------------------------------ projection
class LanguageProjection(_KernelModule):
    """nn.Linear(768, H_llm) (xblip2.py:1552) on the tcgen05 GEMM; accepts [..., 768] fp32/bf16, returns fp32."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features).normal_(0.0, 0.02))
        self.bias = nn.Parameter(torch.zeros(out_features))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._device_check(x)
        if self._packed is None:
            self._packed = (self._bf16(self.weight), self._f32(self.bias))
        w, b = self._packed
        x2 = x.reshape(-1, self.in_features)
        a = x2.contiguous() if x2.dtype == torch.bfloat16 else ops.cast_bf16(x2.to(torch.float32).contiguous())
        y = ops.gemm(a, w, b, ops.EPI_F32)
        return y.view(*x.shape[:-1], self.out_features)


# ------------------------------------------------------------------------------------------------ containers
class _ConditionalGenerationBase(nn.Module):
    instruct = False
    main_input_name = "pixel_values"
    _vision_cls, _qformer_cls = Blip2VisionModel, Blip2QFormerModel
    _hf_config = "Blip2Config"

    def __init__(self, config=None, language_model=None):
        """`config`: an HF `Blip2Config` / `InstructBlipConfig` (or any object with `vision_config`, `qformer_config`,
        `text_config.hidden_size`, `num_query_tokens`, `use_decoder_only_language_model`).  `language_model`: None =
        build the unchanged HF model from `config.text_config` as the reference does; False = none (visual tokens
        only); a module = use it."""
        super().__init__()
        if config is None:
            config = SimpleNamespace(vision_config=None, qformer_config=None, text_config=SimpleNamespace(hidden_size=2048),
                                     num_query_tokens=32, use_decoder_only_language_model=False)
        self.config = config
        vcfg, qcfg, tcfg = _cfg(config, "vision_config"), _cfg(config, "qformer_config"), _cfg(config, "text_config")
        llm_hidden = _cfg(tcfg, "hidden_size", None) or _cfg(tcfg, "d_model", 2048)
        self.vision_model = self._vision_cls(vcfg)
        self.qformer = self._qformer_cls(qcfg)
        self.query_tokens = nn.Parameter(torch.zeros(1, _cfg(config, "num_query_tokens", 32), self.qformer.hidden))
        self.language_projection = LanguageProjection(self.qformer.hidden, llm_hidden)
        self.temporal_projection = _Linear(self.qformer.hidden, llm_hidden)     # checkpoint compatibility only
        if language_model is None:
            import transformers
            auto = transformers.AutoModelForCausalLM if _cfg(config, "use_decoder_only_language_model", False) \
                else transformers.AutoModelForSeq2SeqLM
            language_model = auto.from_config(tcfg)
        self.language_model = language_model if language_model is not False else None

    # -- the accessors the task modules use (LSTP_SF_blip2_module.py:495)
    def get_input_embeddings(self):
        if self.language_model is None:
            raise RuntimeError("this container was built without a language model (language_model=False)")
        return self.language_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.language_model.set_input_embeddings(value)

    def get_output_embeddings(self):
        return self.language_model.get_output_embeddings()

    def get_encoder(self):
        return self.language_model.get_encoder()

    def get_decoder(self):
        return self.language_model.get_decoder()

    def forward(self, *a, **k):
        raise NotImplementedError("the reference task modules never call the container's forward(): they drive "
                                  "vision_model / qformer / language_projection / language_model.generate themselves")

    generate = forward

    # -- construction from checkpoints (LSTP_SF_blip2_module.py:84, eval/utils/model.py:31-32)
    @classmethod
    def from_config(cls, config, **kw):
        return cls(config, **kw)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, config=None, language_model=None,
                        torch_dtype: Optional[torch.dtype] = None, **unused):
        """Reads `config.json` with the HF config class and the weights from the checkpoint directory.  Hot-path
        modules keep fp32 masters (the kernels pack their own bf16 operands); `torch_dtype` applies to the language
        model only.  With `language_model=False` the LLM shards are not read at all."""
        import transformers
        if config is None:
            config = getattr(transformers, cls._hf_config).from_pretrained(pretrained_model_name_or_path)
        model = cls(config, language_model=language_model)
        prefixes = None if model.language_model is not None else (
            "vision_model.", "qformer.", "query_tokens", "language_projection.", "temporal_projection.")
        sd = read_checkpoint(pretrained_model_name_or_path, prefixes)
        if model.language_model is None:
            sd = {k: v for k, v in sd.items() if not k.startswith("language_model.")}
        res = model.load_state_dict({k: v.float() if v.is_floating_point() and not k.startswith("language_model.") else v
                                     for k, v in sd.items()}, strict=False)
        missing = [k for k in res.missing_keys if not k.startswith("temporal_projection.")]     # absent in stock HF checkpoints
        _report(missing, res.unexpected_keys, f"{cls.__name__}.from_pretrained({pretrained_model_name_or_path})")
        if torch_dtype is not None and model.language_model is not None:
            model.language_model.to(torch_dtype)
        model.eval()
        return model


class Blip2ForConditionalGeneration(_ConditionalGenerationBase):
    pass


class InstructBlipForConditionalGeneration(_ConditionalGenerationBase):
    instruct = True
    _vision_cls, _qformer_cls = InstructBlipVisionModel, InstructBlipQFormerModel
    _hf_config = "InstructBlipConfig"
