"""Drop-in for the reference `src/models/components/xinstructblip.py` hot-path classes."""
from videotgb_b200.modeling import InstructBlipQFormerModel, InstructBlipVisionModel  # noqa: F401
