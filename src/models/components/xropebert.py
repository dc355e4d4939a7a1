"""Drop-in for the reference `src/models/components/xropebert.py` temporal sampler."""
from videotgb_b200.sampler import RopeBertModel  # noqa: F401
