"""Overlay package: this directory first, then the reference's `src/models` (task modules live there)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
