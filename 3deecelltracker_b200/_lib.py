"""ctypes binding of libct3d.so (the C ABI declared in include/ct3d.h).

There is deliberately no fallback: if the shared library is missing or a CUDA device is absent, calls
raise.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C 3deecelltracker_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CT3D_LIB") or os.path.join(_HERE, "libct3d.so")     # CT3D_LIB: debug builds of the same library

c_void_p, c_int, c_size_t, c_double, c_float, c_longlong = C.c_void_p, C.c_int, C.c_size_t, C.c_double, C.c_float, C.c_longlong


class CtUNetSpec(C.Structure):
    _fields_ = [("in_x", c_int), ("in_y", c_int), ("in_z", c_int),
                ("pool_x", c_int), ("pool_y", c_int), ("pool_z", c_int),
                ("act_relu", c_int), ("levels", c_int),
                ("down", (c_int * 2) * 4), ("up", (c_int * 2) * 4), ("out", c_int * 2)]


class CtPrglsParams(C.Structure):
    _fields_ = [("mode", c_int), ("max_iteration", c_int), ("beta", c_double), ("lambda_", c_double),
                ("vol", c_double), ("threshold", c_double)]


class CtPrglsProblem(C.Structure):
    _fields_ = [("ref", c_void_p), ("tgt", c_void_p), ("corr", c_void_p), ("tracked", c_void_p),
                ("post", c_void_p), ("ref_out", c_void_p), ("coef", c_void_p), ("tracked_out", c_void_p),
                ("iterations", c_void_p),
                ("n_ref", c_int), ("n_tgt", c_int), ("n_tracked", c_int),
                ("corr_is_f64", c_int), ("prior_given", c_int)]


# symbol -> (restype, argtypes); every symbol declared in include/ct3d.h appears here
SIGNATURES = {
    "ct_abi_version": (c_int, []),
    "ct_last_error": (C.c_char_p, []),
    "ct_launch_count": (C.c_ulonglong, []),
    "ct_set_reserved_sms": (c_int, [c_int]),
    "ct_profile_enable": (c_int, [c_int]),
    "ct_profile_read": (c_int, [c_int, C.POINTER(c_double), C.POINTER(C.c_ulonglong), c_int]),
    "ct_normalize_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ct_normalize_image": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_int, c_int, c_int,
                                   c_void_p, c_size_t, c_void_p]),
    "ct_median": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_select_state_bytes": (c_size_t, []),
    "ct_select_hist_offset": (c_size_t, []),
    "ct_select_passes": (c_int, [c_int]),
    "ct_select_begin": (c_int, [c_void_p, c_longlong, c_void_p]),
    "ct_select_hist": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_int, c_void_p]),
    "ct_select_scan": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "ct_select_finish": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "ct_normalize_image_with_median": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_int, c_int,
                                               c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_unet_weight_count": (c_size_t, [C.POINTER(CtUNetSpec)]),
    "ct_unet_create": (c_int, [C.POINTER(CtUNetSpec), c_void_p, c_size_t, C.POINTER(c_void_p)]),
    "ct_unet_destroy": (None, [c_void_p]),
    "ct_unet_set_engine": (c_int, [c_void_p, c_int]),
    "ct_unet_flops_per_tile": (c_double, [c_void_p]),
    "ct_unet_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "ct_unet_predict_tiles": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "ct_unet_conv_block_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ct_unet_conv_block": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                   c_size_t, c_void_p]),
    "ct_unet_tile_count": (c_int, [c_void_p, c_int, c_int, c_int, C.POINTER(c_int * 3), C.POINTER(c_int * 3)]),
    "ct_unet3_prediction": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, C.POINTER(c_int * 3),
                                    c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "ct_unet3_prediction_block": (c_int, [c_void_p, c_void_p, C.POINTER(c_int * 3), C.POINTER(c_int * 3), c_void_p,
                                          C.POINTER(c_int * 3), C.POINTER(c_int * 3), c_int, c_int, c_int,
                                          C.POINTER(c_int * 3), C.POINTER(c_int * 3), C.POINTER(c_int * 3),
                                          c_void_p, c_size_t, c_int, c_void_p]),
    "ct_ffn_weight_count": (c_size_t, []),
    "ct_ffn_create": (c_int, [c_void_p, c_size_t, C.POINTER(c_void_p)]),
    "ct_ffn_destroy": (None, [c_void_p]),
    "ct_knn_features": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ct_ffn_match_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ct_ffn_match": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_ffn_predict_workspace_bytes": (c_size_t, [c_int]),
    "ct_ffn_predict": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_greedy_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ct_greedy_prior": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "ct_prgls_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ct_prgls": (c_int, [C.POINTER(CtPrglsParams), C.POINTER(CtPrglsProblem), c_int, c_void_p, c_size_t, c_void_p]),
    "ct_predict_one_rep": (c_int, [c_void_p, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    "ct_trim_mean": (c_int, [c_void_p, c_int, c_int, c_double, c_void_p, c_void_p]),
    "ct_replay_fit": (c_int, [c_void_p, c_int, c_int, C.POINTER(c_void_p), C.POINTER(c_int), C.POINTER(c_double),
                            C.POINTER(c_void_p), c_double, c_void_p, c_void_p, c_void_p]),
    "ct_watershed_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ct_label_components_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ct_label_components": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_correction_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "ct_accurate_correction": (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_int, c_int, c_int, c_int, c_double] + [c_void_p] * 5 + [c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_tracked_labels_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "ct_tracked_labels": (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_recalculate_cell_boundaries_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ct_recalculate_cell_boundaries": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ct_watershed_segment": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class Ct3dError(RuntimeError):
    pass


def lib():
    """Load libct3d.so once.  Raises ImportError when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                              "(or __graft_entry__.build()); this package has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.ct_abi_version() != 1:
            raise ImportError("libct3d.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc, exc=Ct3dError):
    if rc != 0:
        raise exc(lib().ct_last_error().decode("utf-8", "replace"))
