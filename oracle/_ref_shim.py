"""Import shim that loads the *unmodified* reference NumPy modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (in the build container, where
/root/reference is mounted) to produce the golden vectors under ``tests/golden``.  Nothing in the
product package, bench.py's GPU arm, or the ``-m gpu`` tests imports this file: /root/reference
does not exist on the GPU box.

The reference modules ``CellTracker/track.py`` and ``CellTracker/trackerlite.py`` are pure
NumPy/SciPy/scikit-learn apart from plotting imports and TF-dependent sibling imports
(track.py:6; trackerlite.py:10,15-17).  Those are replaced by empty stub modules so the numerical
functions run verbatim.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CT3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "CellTracker", "track.py"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def load():
    """Return (track_module, trackerlite_module) of the reference, imported verbatim."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k.startswith("matplotlib") or k.startswith("CellTracker")}
    try:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.patches = _stub("matplotlib.patches", ConnectionPatch=object)
        pkg = _stub("CellTracker")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "CellTracker")]
        _stub("CellTracker.coord_image_transformer", Coordinates=object,
              plot_prgls_prediction=None, plot_two_pointset_scatters=None)
        _stub("CellTracker.ffn", initial_matching_ffn=None, normalize_points=None, FFN=object)
        _stub("CellTracker.stardistwrapper", load_2d_slices_at_time=None)
        track = importlib.import_module("CellTracker.track")
        lite = importlib.import_module("CellTracker.trackerlite")
        return track, lite
    finally:
        for k in [k for k in sys.modules if k.startswith("matplotlib") or k.startswith("CellTracker")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
