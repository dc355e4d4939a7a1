"""Stand-ins for the third-party packages the reference task modules import at module level but this image lacks
(lightning, torchmetrics, sentence_transformers, peft).  Only names are needed: nothing on the hot path calls them.
Test infrastructure for tests/test_boundary_*.py."""
import importlib.machinery
import sys
import types


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Anything()


_NAMES = {
    "lightning": ["LightningModule"],
    "torchmetrics": ["MaxMetric", "MeanMetric"],
    "torchmetrics.classification": [],
    "torchmetrics.classification.accuracy": ["Accuracy"],
    "torchmetrics.text": [],
    "torchmetrics.text.bleu": ["BLEUScore"],
    "torchmetrics.text.bert": ["BERTScore"],
    "torchmetrics.text.rouge": ["ROUGEScore"],
    "torchmetrics.text.perplexity": ["Perplexity"],
    "sentence_transformers": ["SentenceTransformer", "util"],
}


def _module_getattr(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Anything


def install():
    import torch
    for name, attrs in _NAMES.items():
        try:
            __import__(name)
            continue
        except ImportError:
            pass
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        for a in attrs:
            setattr(m, a, _Anything)
        m.__getattr__ = _module_getattr          # any other name (torchmetrics.Metric, ...) is a harmless base class
        sys.modules[name] = m
    lt = sys.modules["lightning"]
    if getattr(lt, "LightningModule", None) is _Anything:
        lt.LightningModule = torch.nn.Module
