"""TEST INFRASTRUCTURE ONLY.  Import the reference's own hot-path modules.

Works only where /root/reference is mounted (this build container; never the GPU
box).  ``src/__init__.py`` pulls in skimage (absent), so the sub-package
``feature_extraction`` is imported with ``/root/reference/src`` on ``sys.path``
(the same trick the reference uses in ``scripts/preprocess_dataset.py:18``) and
``training/losses.py`` is loaded by file path.  The only substituted piece is the
third-party ``efficientnet_pytorch`` (oracle/efficientnet_pytorch).
"""
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import types

from . import ensure_shim_on_path

REFERENCE_DIR = os.environ.get("DFV_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "src", "feature_extraction", "feature_extractor.py"))


_cache = None


def load_reference() -> types.SimpleNamespace:
    """Returns a namespace with the reference classes (verbatim reference code)."""
    global _cache
    if _cache is not None:
        return _cache
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_DIR}")
    ensure_shim_on_path()
    src = os.path.join(REFERENCE_DIR, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    fe = importlib.import_module("feature_extraction")
    spec = importlib.util.spec_from_file_location(
        "dfv_reference_losses", os.path.join(src, "training", "losses.py"))
    losses = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(losses)
    _cache = types.SimpleNamespace(
        DeepfakeDetectionModel=fe.DeepfakeDetectionModel,
        DeepfakeFeatureExtractor=fe.DeepfakeFeatureExtractor,
        EfficientNetB4Backbone=fe.EfficientNetB4Backbone,
        HybridAttention=fe.HybridAttention,
        LandmarkAttention=fe.LandmarkAttention,
        ChannelAttention=fe.ChannelAttention,
        SpatialAttention=fe.SpatialAttention,
        CombinedLoss=losses.CombinedLoss,
        FocalLoss=losses.FocalLoss,
        ContrastiveLoss=losses.ContrastiveLoss,
        kind="reference",
    )
    return _cache


@contextlib.contextmanager
def quiet():
    """The reference prints status lines from its constructors; silence them."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
