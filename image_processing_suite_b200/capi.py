"""ctypes binding of the C ABI declared in include/ips.h (``libips.so``).

This is the only place the shared library is loaded.  There is no fallback: if the
library has not been built (``python -m image_processing_suite_b200.build``) loading
raises, and every compute entry point returns IPS_ERR_CUDA -> ``IpsError`` when no CUDA
device is present.
"""
import ctypes as C
import os
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libips.so")

IPS_OK = 0
ERR_NAMES = {-1: "IPS_ERR_BAD_SHAPE", -2: "IPS_ERR_BAD_DTYPE", -3: "IPS_ERR_BAD_ALIGN",
             -4: "IPS_ERR_CUDA", -5: "IPS_ERR_NCCL", -6: "IPS_ERR_NOMEM", -7: "IPS_ERR_BAD_ARG"}


class IpsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "IPS_ERR"), code, message))
        self.code = code


p = C.c_void_p
i = C.c_int
f = C.c_float
sz = C.c_size_t
u64 = C.c_uint64
i64 = C.c_int64

# name -> (restype, argtypes); status-returning functions use restype int and are checked.
PROTOTYPES = {
    "ips_abi_version": (i, []),
    "ips_last_error": (C.c_char_p, []),
    "ips_launch_count": (u64, []),
    "ips_device_info": (i, [C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(sz)]),
    "ips_preprocess_workspace_bytes": (sz, [i, i, i, i, i]),
    "ips_preprocess_fused": (i, [p, p, p, p, p, i, p, p, sz, i, i, i, i, i, p]),
    "ips_object_stats_workspace_bytes": (sz, [i, i, i]),
    "ips_object_stats": (i, [p, p, p, f, p, p, p, i, p, sz, i, i, i, i, p]),
    "ips_field_fused_workspace_bytes": (sz, [i, i, i, i, i, i]),
    "ips_field_fused": (i, [p, p, p, p, p, i, f, p, p, p, i, p, sz, i, i, i, i, i, p]),
    "ips_field_fused_ex": (i, [p, p, i, p, i, p, p, i, f, p, p, p, i, p, sz, i, i, i, i, i, p]),
    "ips_illum_reciprocal": (i, [p, p, i64, p]),
    "ips_illum_accumulate": (i, [p, p, i, i, i, i, p]),
    "ips_illum_finalize_workspace_bytes": (sz, [i, i, i]),
    "ips_illum_finalize": (i, [p, u64, C.c_double, C.c_double, p, p, sz, i, i, i, p]),
    "ips_illum_median": (i, [p, p, i, i, i, i, p]),
    "ips_illum_smooth_rescale": (i, [p, C.c_double, C.c_double, p, p, sz, i, i, i, p]),
    "ips_lanczos_workspace_bytes": (sz, [i, i, i, i, i]),
    "ips_lanczos_resize_u16": (i, [p, p, i, i, i, i, i, p, sz, p]),
    "ips_ring_sums": (i, [p, p, p, i, i, i, i, p]),
    "ips_rps_prepare_workspace_bytes": (sz, [i64]),
    "ips_rps_prepare": (i, [p, p, p, p, p, i64, p, sz, p]),
    "ips_ring_sums_half": (i, [p, p, p, i, i, i, i, p]),
    "ips_loglog_slope": (i, [p, p, i, i, p]),
    "ips_cosine_workspace_bytes": (sz, [i, i]),
    "ips_cosine_triu": (i, [p, p, i, p, p, i, i, p, sz, p]),
    "ips_cosine_planes_bytes": (sz, [i, i]),
    "ips_cosine_plane_row_bytes": (sz, [i]),
    "ips_cosine_plane_stride_bytes": (sz, [i, i]),
    "ips_cosine_split_rows": (i, [p, i, i, i, i, p, p]),
    "ips_cosine_triu_part": (i, [p, p, i, i, i, i, p]),
    "ips_cosine_pairs_workspace_bytes": (sz, [i, i]),
    "ips_cosine_triu_pairs": (i, [p, p, i, p, p, p, p, u64, i, i, p, sz, p]),
    "ips_well_mean_workspace_bytes": (sz, [i, i]),
    "ips_well_mean": (i, [p, p, p, p, i, i, i, p, sz, p]),
    "ips_well_sums_reset": (i, [p, sz, i, i, p]),
    "ips_well_sums_add": (i, [p, p, i64, p, sz, i, i, p]),
    "ips_well_sums_finalize": (i, [p, sz, p, p, i, i, p]),
    "ips_well_sums_add_blocks": (i, [p, i64, i64, p, sz, i, i, p]),
    "ips_well_mean_f64": (i, [p, p, p, p, i64, i, i, p, sz, p]),
    "ips_well_median_f64": (i, [p, p, p, p, p, i64, i, i, p]),
    "ips_cell_crops_workspace_bytes": (sz, [i, i]),
    "ips_cell_crops": (i, [p, p, p, p, i, i, p, p, p, p, sz, i, i, i, i, i, p]),
    "ips_tiff_rows_per_strip": (i, [i, i]),
    "ips_tiff_lzw_bound": (sz, [sz]),
    "ips_tiff_file_bound": (sz, [i, i, i]),
    "ips_tiff_encode_workspace_bytes": (sz, [i, i, i, i]),
    "ips_tiff_lzw_encode_u16": (i, [p, i, i, i, i, p, sz, p, p, sz, p]),
    "ips_tiff_lzw_decode": (i, [p, p, p, p, p, p, i, p, p]),
    "ips_tiff_fix_u16": (i, [p, i64, i, i, i, p]),
    "ips_mad_robustize": (i, [p, p, p, i, i, p]),
    "ips_double_sigmoid_abs": (i, [p, p, i64, i, C.c_double, p]),
    "ips_pack_rows_workspace_bytes": (sz, [i]),
    "ips_pack_rows": (i, [p, p, p, p, i, p, p, i, i, i, p, sz, p]),
    "ips_pack_rows_block": (i, [p, p, p, p, i, p, i64, i, i, i, p, sz, p]),
    "ips_block_counts": (i, [p, p, i64, i64, i, p]),
    "ips_rows_well_ids": (i, [p, p, p, i64, i, i, p]),
    "ips_comm_unique_id_bytes": (i, []),
    "ips_comm_unique_id": (i, [p, i]),
    "ips_comm_create": (i, [C.POINTER(p), p, i, i, i]),
    "ips_comm_destroy": (i, [p]),
    "ips_allgather_rows": (i, [p, p, i64, i, p, p, i64, p]),
    "ips_allgather_blocks": (i, [p, p, i64, i, p]),
    "ips_comm_rank": (i, [p, C.POINTER(i), C.POINTER(i)]),
    "ips_device_alloc": (i, [C.POINTER(p), sz]),
    "ips_device_free": (i, [p]),
    "ips_ipc_handle_bytes": (i, []),
    "ips_ipc_export": (i, [p, p, i]),
    "ips_peer_table_open": (i, [C.POINTER(p), p, i, i, p]),
    "ips_peer_table_close": (i, [p]),
    "ips_peer_push": (i, [p, sz, sz, p]),
    "ips_host_alloc": (i, [C.POINTER(p), sz]),
    "ips_host_alloc_flags": (i, [C.POINTER(p), sz, C.c_uint]),
    "ips_host_free": (i, [p]),
    "ips_pipeline_create": (i, [C.POINTER(p), i, i, i, i, i, i, i, i, i, p, f]),
    "ips_pipeline_submit": (i64, [p, p, p, p, p, p, p, p]),
    "ips_pipeline_wait": (i, [p, i64]),
    "ips_pipeline_destroy": (i, [p]),
}
UNCHECKED = {"ips_abi_version", "ips_last_error", "ips_launch_count"}

_lib = None
_lock = threading.Lock()


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        "libips.so is missing (%s): build it with `python -m image_processing_suite_b200.build`; "
                        "there is no CPU fallback" % LIB_PATH)
                _lib = C.CDLL(LIB_PATH)
    return _lib


_bound = {}


def fn(name):
    """The raw ctypes function with restype / argtypes set."""
    g = _bound.get(name)
    if g is None:
        res, args = PROTOTYPES[name]
        g = getattr(lib(), name)
        g.restype = res
        g.argtypes = args
        _bound[name] = g
    return g


def last_error():
    return fn("ips_last_error")().decode("utf-8", "replace")


def call(name, *args):
    """Call an entry point; raise IpsError on a negative status."""
    r = fn(name)(*args)
    if name not in UNCHECKED and PROTOTYPES[name][0] in (i, i64) and r < 0:
        raise IpsError(int(r), last_error())
    return r


def launch_count():
    return int(fn("ips_launch_count")())
