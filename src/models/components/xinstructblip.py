"""Drop-in for the reference `src/models/components/xinstructblip.py` (`InstructBlipForConditionalGeneration`,
src/models/LSTP_SF_module.py:19) and the hot-path classes underneath."""
from videotgb_b200.containers import InstructBlipForConditionalGeneration  # noqa: F401
from videotgb_b200.modeling import InstructBlipQFormerModel, InstructBlipVisionModel  # noqa: F401
