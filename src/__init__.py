"""Overlay of the reference's `src` package.

Put this repo BEFORE the reference checkout on `sys.path` / `PYTHONPATH`.  `src`, `src.models` and
`src.models.components` then resolve here first and fall through to the reference's own directories for everything this
repo does not provide (`src.gadgets`, `src.data`, `src.models.LSTP_*`, `src.models.components.xsampler`, ...), so the
reference's task modules import unchanged and pick up the B200 implementations of the hot-path classes through the
import statements they already have (`from .components.xblip2 import Blip2ForConditionalGeneration`, ...,
src/models/LSTP_SF_blip2_module.py:20-25)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
