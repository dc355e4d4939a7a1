"""Drop-in for the reference `src/models/components/xblip2.py` hot-path classes (sm_100a kernels underneath)."""
from videotgb_b200.modeling import Blip2QFormerModel, Blip2VisionModel  # noqa: F401
from videotgb_b200.pipeline import LanguageProjection  # noqa: F401
