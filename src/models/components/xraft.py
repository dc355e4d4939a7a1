"""Drop-in for the reference `src/models/components/xraft.py` (`from .components.xraft import RAFT, InputPadder`,
src/models/LSTP_SF_blip2_module.py:24): the optical-flow extractor on the sm_100a kernels."""
from videotgb_b200.raft import RAFT, InputPadder  # noqa: F401
