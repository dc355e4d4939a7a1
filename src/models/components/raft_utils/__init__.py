"""Overlay of the reference's `raft_utils` directory: `utils` is provided here, the rest falls through."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
