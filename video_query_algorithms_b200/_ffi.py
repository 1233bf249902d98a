"""ctypes binding of include/vq.h.  There is no CPU fallback: if the CUDA library is missing
or a call fails, VQError is raised (the broker's `except` logs it, reference src/broker.py:88-89)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VQ_B200_LIB: development override (e.g. a build with role cycle counters); the default is the in-tree build
LIB_PATH = os.environ.get("VQ_B200_LIB") or os.path.join(HERE, "lib", "libvq_b200.so")
MAX_STREAMS = 4
MAX_TOPK = 1024


class VQError(RuntimeError):
    pass


class ScanParams(C.Structure):
    _fields_ = [("weights", C.c_double * MAX_STREAMS), ("threshold", C.c_double),
                ("lower_limit", C.c_double), ("eps", C.c_double), ("topk", C.c_int32),
                ("want_sims", C.c_int32)]


class ScanCounts(C.Structure):
    _fields_ = [("n_match", C.c_int64), ("n_near", C.c_int64), ("n_tie", C.c_int64),
                ("n_topk", C.c_int32), ("scan_ms", C.c_float)]


class ScanDeviceView(C.Structure):
    _fields_ = [("scores_dev", C.c_void_p), ("sims_dev", C.c_void_p), ("counts_dev", C.c_void_p),
                ("topk_scores_dev", C.c_void_p), ("topk_rows_dev", C.c_void_p),
                ("match_rows_dev", C.c_void_p), ("near_rows_dev", C.c_void_p),
                ("match_scores_dev", C.c_void_p), ("near_scores_dev", C.c_void_p),
                ("tie_rows_dev", C.c_void_p), ("tie_scores_dev", C.c_void_p)]


_P = C.POINTER
_vp, _i32, _i64, _f32p, _f64p, _i64p = C.c_void_p, C.c_int32, C.c_int64, _P(C.c_float), _P(C.c_double), _P(C.c_int64)

# name -> (restype, argtypes); every symbol include/vq.h declares
PROTOTYPES = {
    "vq_last_error": (C.c_char_p, []),
    "vq_abi_version": (C.c_int, []),
    "vq_device_count": (C.c_int, [_P(C.c_int)]),
    "vq_store_create": (C.c_int, [_P(_vp), C.c_int, _i64, C.c_int, C.c_int, C.c_int, _i64]),
    "vq_store_destroy": (C.c_int, [_vp]),
    "vq_store_reserve": (C.c_int, [_vp, _i64]),
    "vq_store_append": (C.c_int, [_vp, _i64, _vp]),
    "vq_store_describe": (C.c_int, [_vp, _i64p, _P(C.c_int), _P(C.c_int), _P(C.c_int), _i64p, _P(C.c_int)]),
    "vq_store_upload": (C.c_int, [_vp, _i64, _i64, _vp]),
    "vq_store_upload_async": (C.c_int, [_vp, _i64, _i64, _vp]),
    "vq_store_sync": (C.c_int, [_vp]),
    "vq_pinned_alloc": (C.c_int, [_P(_vp), _i64]),
    "vq_pinned_free": (C.c_int, [_vp]),
    "vq_store_download": (C.c_int, [_vp, _i64, _i64, _vp]),
    "vq_store_set_split_weights": (C.c_int, [_vp, _vp]),
    "vq_store_fill_synthetic": (C.c_int, [_vp, C.c_uint64, _vp]),
    "vq_store_device_ptr": (C.c_int, [_vp, _P(_vp)]),
    "vq_scan": (C.c_int, [_vp, _vp, _P(ScanParams), _P(ScanCounts)]),
    "vq_scan_select": (C.c_int, [_vp, _vp, _P(ScanParams), _P(ScanCounts), _i64p, _i64p, _f32p]),
    "vq_gather_list": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp]),
    "vq_scan_multi": (C.c_int, [_vp, _i32, _vp, _P(ScanParams), _i32, _vp, _vp, _i32, _vp, _vp, _P(_i32)]),
    "vq_gather_list_multi": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "vq_scan_multi_host_list": (C.c_int, [_vp, _i32, _P(_vp), _P(_vp), _i64p]),
    "vq_fetch_scores_at": (C.c_int, [_vp, _i64, _vp, _vp]),
    "vq_scan_phase_times": (C.c_int, [_vp, _i32, _vp, _vp, _P(_i32)]),
    "vq_exchange_check": (C.c_int, [_vp]),
    "vq_exchange_kernel_times": (C.c_int, [_vp, _i32, _vp, _vp, _P(_i32)]),
    "vq_rank_list": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp]),
    "vq_mt_sample_range": (C.c_int, [_vp, _P(_i32), _i64, _i64, _vp]),
    "vq_scan_enqueue": (C.c_int, [_vp, _vp, _P(ScanParams), _vp]),
    "vq_scan_wait": (C.c_int, [_vp, _vp, _P(ScanCounts)]),
    "vq_fetch_matches": (C.c_int, [_vp, _i64, _vp, _vp]),
    "vq_fetch_near": (C.c_int, [_vp, _i64, _vp, _vp]),
    "vq_fetch_ties": (C.c_int, [_vp, _i64, _vp, _vp]),
    "vq_fetch_topk": (C.c_int, [_vp, _i32, _vp, _vp]),
    "vq_scan_host_list": (C.c_int, [_vp, _i32, _P(_vp), _P(_vp), _i64p]),
    "vq_fetch_ranked": (C.c_int, [_vp, _i32, _i64, _vp, _vp]),
    "vq_fetch_scores": (C.c_int, [_vp, _i64, _i64, _vp]),
    "vq_fetch_sims": (C.c_int, [_vp, _i64, _i64, _vp]),
    "vq_scan_view": (C.c_int, [_vp, _P(ScanDeviceView)]),
    "vq_scan_payload": (C.c_int, [_vp, _P(_vp), _P(_i32)]),
    "vq_merge_payloads_enqueue": (C.c_int, [C.c_int, _vp, _i32, _i32, _vp, _vp]),
    "vq_exchange_create": (C.c_int, [_P(_vp), C.c_int, C.c_int, C.c_int]),
    "vq_exchange_local_handle": (C.c_int, [_vp, _vp]),
    "vq_exchange_connect": (C.c_int, [_vp, _vp]),
    "vq_exchange_connect_local": (C.c_int, [_vp, _i32]),
    "vq_exchange_destroy": (C.c_int, [_vp]),
    "vq_scan_exchange_enqueue": (C.c_int, [_vp, _vp, _vp]),
    "vq_scan_exchange_enqueue_lagged": (C.c_int, [_vp, _vp, _vp]),
    "vq_exchange_flush_enqueue": (C.c_int, [_vp, _vp]),
    "vq_exchange_merged": (C.c_int, [_vp, _P(_vp)]),
    "vq_hostx_create": (C.c_int, [_P(_vp), C.c_char_p, C.c_int, C.c_int, _i64]),
    "vq_hostx_unlink": (C.c_int, [_vp]),
    "vq_hostx_allgather": (C.c_int, [_vp, _vp, _i64, _vp, C.c_double]),
    "vq_hostx_destroy": (C.c_int, [_vp]),
    "vq_summary_pack": (C.c_int, [_i64, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vq_summary_merge": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _P(_i32), _vp, _vp, _vp, _i64p]),
    "vq_scan_kernel_times": (C.c_int, [_vp, _i32, _vp, _P(_i32)]),
    "vq_merge_topk": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _P(_i32)]),
    "vq_merge_topk_batch": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "vq_labelled_sims": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "vq_loss_grid": (C.c_int, [C.c_int, _vp, _vp, _i64, _vp, _i32, _vp, _i32, C.c_double, _vp, _vp, _i32, _vp]),
    "vq_bootstrap_target": (C.c_int, [_vp, _vp, _i32, _vp, _i32, C.c_double, _vp, _vp]),
    "vq_scan_batch": (C.c_int, [_vp, _vp, _i32, _P(ScanParams), _vp, _vp, _vp, _vp]),
    "vq_scan_batch_ties": (C.c_int, [_vp, _vp, _i32, _P(ScanParams), _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "vq_scan_batch_scores": (C.c_int, [_vp, _vp, _i32, _P(ScanParams), _vp]),
    "vq_csv_shape": (C.c_int, [C.c_char_p, _i64p, _P(_i32), C.c_char_p, _i32]),
    "vq_csv_read": (C.c_int, [C.c_char_p, _i32, _i64, _i32, _vp, _vp, _i64p]),
}

_lib = None


def lib():
    """The loaded library; raises VQError if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VQError("libvq_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(expected at %s)" % LIB_PATH)
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:
            raise VQError("cannot load %s: %s" % (LIB_PATH, e))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


PYHOST_PATH = os.path.join(HERE, "lib", "libvq_pyhost.so")
_pyhost = None


def pyhost():
    """libvq_pyhost.so (csrc/vq_pyhost.c): unboxes API records into the store's row layout.  PyDLL: its functions take
    Python objects and are called with the GIL held."""
    global _pyhost
    if _pyhost is None:
        if not os.path.exists(PYHOST_PATH):
            raise VQError("libvq_pyhost.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        h = C.PyDLL(PYHOST_PATH)
        h.vq_py_last_error.restype = C.c_char_p
        h.vq_py_index_records.restype = C.c_int
        h.vq_py_index_records.argtypes = [C.py_object, C.py_object, C.py_object, _vp, _vp, _vp, _vp]
        h.vq_py_fill_chunk.restype = C.c_int
        h.vq_py_fill_chunk.argtypes = [C.py_object, _vp, _vp, _i64, _i64, _vp, _i32]
        _pyhost = h
    return _pyhost


def check_py(rc, what):
    if rc != 0:
        raise VQError("%s failed: %s" % (what, (pyhost().vq_py_last_error() or b"?").decode()))


def check(rc, what=""):
    if rc != 0:
        msg = lib().vq_last_error()
        raise VQError("%s failed (%d): %s" % (what or "libvq_b200 call", rc, msg.decode() if msg else "?"))


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags.c_contiguous
    return C.c_void_p(a.ctypes.data)          # (`data_as` costs twice as much, and this sits on the per-query path)
