"""CPU: the C-ABI library builds, loads and exports exactly what include/videotgb_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from videotgb_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared():
    with open(os.path.join(ROOT, "include", "videotgb_b200.h")) as f:
        return sorted(set(re.findall(r"TGB_API\s+[\w\s\*]+?\b(tgb_\w+)\s*\(", f.read())))


def test_header_symbols_exported(lib):
    from videotgb_b200 import _lib
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string(lib):
    assert lib.tgb_version() >= 100
    assert isinstance(lib.tgb_last_error(), bytes)


def test_no_fallback_without_gpu():
    import torch
    from videotgb_b200 import ops, _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.TgbError):
        ops.normalize_frames_u8(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "videotgb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
