"""Reference import paths for the hot-path components, backed by the sm_100a kernels; modules this repo does not
replace (`xsampler`, `xvit`, `xropealbert`, `raft_utils.{corr,extractor,update}`) fall through to the reference."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
