"""Import shim that lets the UNMODIFIED reference components import under transformers 5.x.

Only used by ``oracle/gen_golden.py`` in the build container (``/root/reference`` does
not exist on the GPU box).  The reference pins transformers 4.36; three helper symbols
moved/disappeared since (`xblip2.py:34`, `xropebert.py:45-50`), `get_head_mask`
(`xblip2.py:1146`) and the bare `init_weights()` call (`xropebert.py:960`) need compat.
None of these touch the arithmetic of the hot path.
"""
import sys

REFERENCE_ROOT = "/root/reference"


def install():
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    from transformers.modeling_utils import PreTrainedModel

    def _unused(*a, **k):            # head pruning is never reached on the hot path
        raise NotImplementedError

    if not hasattr(pu, "find_pruneable_heads_and_indices"):
        pu.find_pruneable_heads_and_indices = _unused
    for n in ("apply_chunking_to_forward", "find_pruneable_heads_and_indices", "prune_linear_layer"):
        if not hasattr(mu, n):
            setattr(mu, n, getattr(pu, n))
    PreTrainedModel.get_head_mask = lambda self, head_mask, n, chunked=False: [None] * n
    if not getattr(PreTrainedModel, "_tgb_shim", False):
        orig = PreTrainedModel.init_weights

        def init_weights_compat(self):
            return self.post_init() if not hasattr(self, "all_tied_weights_keys") else orig(self)

        PreTrainedModel.init_weights = init_weights_compat
        PreTrainedModel._tgb_shim = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
