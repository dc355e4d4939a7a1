"""Drop-in for the reference `src/models/components/xblip2.py`: the names its task modules import
(`Blip2ForConditionalGeneration`, src/models/LSTP_SF_blip2_module.py:21) and the hot-path classes underneath."""
from videotgb_b200.containers import Blip2ForConditionalGeneration, LanguageProjection  # noqa: F401
from videotgb_b200.modeling import Blip2QFormerModel, Blip2VisionModel  # noqa: F401
