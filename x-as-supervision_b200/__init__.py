"""xsup_b200 — B200-native integral (soft-argmax) multi-hypothesis head fused with the
per-hypothesis perspective geometry and the min-over-hypotheses loss of X-as-Supervision.

The directory is named `x-as-supervision_b200`; import it as `xsup_b200` (the repository root
holds a one-line alias module) or with `importlib.import_module("x-as-supervision_b200")`.

Sub-modules that touch the GPU (`ops`, `detector`, `skeleton`, `evalops`, `losses`, `model`) load `csrc/libxsup_b200.so` on import and
raise if it is missing — there is no CPU or PyTorch fallback.  `synth` and `dist` are pure host
code and import anywhere.
"""
from . import dist, synth  # noqa: F401

__all__ = ["dist", "synth", "ops", "detector", "skeleton", "evalops", "losses", "model", "load_native"]


def load_native():
    """Import the CUDA-backed sub-modules (needs the built shared library)."""
    from . import detector, evalops, losses, model, ops, skeleton  # noqa: F401
    return ops


def __getattr__(name):
    if name in ("ops", "detector", "skeleton", "evalops", "losses", "model", "_cabi"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
