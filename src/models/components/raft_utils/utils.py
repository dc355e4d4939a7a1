"""`from .components.raft_utils.utils import dp_state_to_normal` (src/models/LSTP_SF_blip2_module.py:25)."""
from videotgb_b200.raft import InputPadder, dp_state_to_normal  # noqa: F401
