"""B200-native (sm_100a) hot path of a 3DeeCellTracker-compatible segment-and-track engine.

The package mirrors the reference's module names for the one accelerated path:

    unet3d        unet3_a / unet3_b / unet3_c, unet3_prediction            (CellTracker/unet3d.py)
    preprocess    _normalize_image, lcn_gpu                                 (CellTracker/preprocess.py)
    ffn           FFN, initial_matching_ffn, normalize_points               (CellTracker/ffn.py)
    track         pr_gls_quick, initial_matching_quick, get_reference_vols  (CellTracker/track.py)
    trackerlite   TrackerLite, simple_match, prgls_with_two_ref, ...        (CellTracker/trackerlite.py)
    tracker       Tracker (segment / match / track loop)                    (CellTracker/tracker.py)

All arithmetic runs in hand-written CUDA behind the C ABI of include/ct3d.h (libct3d.so); Python holds
PyTorch tensors as device containers and NumPy arrays at the reference-facing boundary.
"""
from . import _lib  # noqa: F401

__all__ = ["unet3d", "preprocess", "ffn", "track", "trackerlite", "tracker", "coord_image_transformer", "dist"]
