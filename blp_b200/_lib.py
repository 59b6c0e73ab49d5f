"""ctypes binding of libblp_b200.so (the C ABI declared in include/blp_b200.h).

The library is the product: there is no Python / PyTorch fallback.  If it has not
been built, or the tensors are not on an sm_100 device, the callers raise.
Build it with `python -m blp_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BLP_B200_LIB") or os.path.join(_HERE, "libblp_b200.so")    # BLP_B200_LIB: tuning builds

MODELS = {"transe": 0, "distmult": 1, "complex": 2, "simple": 3}
LOSSES = {"margin": 0, "nll": 1}

# every symbol include/blp_b200.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
SYMBOLS = {
    "blp_version": (_i32, []),
    "blp_last_error": (ctypes.c_char_p, []),
    "blp_device_check": (_i32, [_i32]),
    "blp_last_launch_count": (_i32, []),
    "blp_score_bcast": (_i32, [_i32, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _vp]),
    "blp_rank_counts": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "blp_metrics_from_counts": (_i32, [_vp, _vp, _i64, ctypes.POINTER(_i64), _i32, _vp, _vp, _vp]),
    "blp_metrics_reduce": (_i32, [_vp, _vp, _i64, ctypes.POINTER(_i64), _i32, _vp, _vp]),
    "blp_eval_rank": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "blp_rank_sweep": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64,
                              _vp, _vp, _vp, _vp, _vp, _vp]),
    "blp_true_scores": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "blp_rank_sweep_counts": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64,
                                     _vp, _vp, _vp, _vp, _vp, _vp]),
    "blp_rank_step_workspace_bytes": (_i64, [_i64]),
    "blp_rank_step": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp, _vp,
                             ctypes.POINTER(_i64), _i32, _vp, _vp, _vp, _vp, _vp]),
    "blp_plan_create": (_i32, [ctypes.POINTER(_vp), _i32, _vp, _i64, _i64, _i32, _vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp,
                               ctypes.POINTER(_i64), _i32, _vp, _vp, _vp, _vp]),
    "blp_plan_run": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "blp_plan_set_overlap": (_i32, [_vp, _i32]),
    "blp_plan_destroy": (None, [_vp]),
    "blp_rank_queries": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp,
                                ctypes.POINTER(_i64), _i32, _vp, _vp, _vp, _vp, _vp]),
    "blp_fast_table_bytes": (_i64, [_i64]),
    "blp_fast_query_bytes": (_i64, [_i64]),
    "blp_fast_prepare_table": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "blp_rank_sweep_fast": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "blp_fast_refine_bytes": (_i64, [_i64]),
    "blp_rank_sweep_fast_exact": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "blp_rank_metrics": (_i32, [_vp, _vp, _i64, ctypes.POINTER(_i64), _i32, _vp, _vp, _vp, _vp]),
    "blp_train_workspace_bytes": (_i64, [_i64, _i64]),
    "blp_train_loss": (_i32, [_i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i64, _i32, _f32,
                              _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "blp_scale": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "blp_pair_loss": (_i32, [_i32, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "blp_l2_regularization": (_i32, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    "blp_filter_index_bytes": (_i64, [_i64]),
    "blp_filter_index_build": (_i32, [_vp, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "blp_filter_correct": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i64,
                                  _vp, _vp, _vp, _vp, _vp, _vp]),
    "blp_filter_correct_rows": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    "blp_filter_correct_mask": (_i32, [_i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    "blp_mrr_breakdown": (_i32, [_vp, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "blp_negative_sample": (_i32, [_i64, _i64, _i64, ctypes.c_uint64, ctypes.c_uint64, _vp, _vp]),
    "blp_store_rows": (_i32, [_vp, _i64, _i32, _i32, _vp, _i64, _vp, _i64, _i64, _vp]),
    "blp_profile_events": (_i32, [_i32, _vp, _vp]),
    "blp_debug_timestamps": (_i32, [_vp]),
    "blp_pipe_probe": (_i32, [_i32, _vp, _i64, _i32, ctypes.POINTER(ctypes.c_double), _vp]),
    "blp_atomic_probe": (_i32, [_vp, _i64, _i32, ctypes.POINTER(ctypes.c_double), _vp]),
}

_lib = None
_lock = threading.Lock()


class BlpError(RuntimeError):
    pass


def lib():
    """Load libblp_b200.so once; raise loudly when it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(SO_PATH):
                    raise BlpError(
                        f"{SO_PATH} is not built; run `python -m blp_b200.build` (needs nvcc). "
                        "blp_b200 has no CPU / PyTorch fallback.")
                handle = ctypes.CDLL(SO_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc, what):
    """Turn a negative BLP_E* return code into the exception the reference would raise."""
    if rc == 0:
        return
    msg = lib().blp_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg} (rc={rc})"
    if rc in (-1, -2):          # BLP_EINVAL / BLP_EDIM: the reference raises ValueError (models.py:26,36)
        raise ValueError(text)
    raise BlpError(text)


def last_launch_count():
    return int(lib().blp_last_launch_count())
