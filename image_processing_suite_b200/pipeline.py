"""Host-buffer pipeline (ips_pipeline_* of include/ips.h): the end-to-end call.

``FieldPipeline.submit`` takes page-locked host arrays of one batch of fields (raw z-stacks
+ label masks) and host output arrays; copies, K1 and K3 run asynchronously on the
library's three streams.  ``pinned_empty`` hands out page-locked NumPy arrays.
"""
import ctypes as C

import numpy as np

from . import capi


class _Pinned:
    """Owns one cudaHostAlloc block; arrays created on it keep it alive via .base."""

    def __init__(self, nbytes, flags=0):
        self.ptr = C.c_void_p()
        capi.call("ips_host_alloc_flags", C.byref(self.ptr), max(int(nbytes), 1), int(flags))
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                capi.fn("ips_host_free")(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype, flags=0):
    """Page-locked host array (uninitialised); the allocation lives as long as the array.
    flags: 1 = portable, 2 = write-combined (ips_host_alloc_flags)."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    blk = _Pinned(count * dtype.itemsize, flags)
    buf = (C.c_char * max(blk.nbytes, 1)).from_address(blk.ptr.value)
    buf._ips_block = blk                      # arr.base -> buf -> blk keeps the memory alive
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def _hp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class FieldPipeline:
    """depth-slot H2D -> K1 -> K3 -> D2H pipeline for batches of ``fields_per_batch`` fields."""

    def __init__(self, fields_per_batch, C_, Z, H, W, bin=2, n_max=2048, depth=3, illum=None,
                 intensity_scale=1.0, label_dtype=np.int32):
        self.shape = (int(fields_per_batch), int(C_), int(Z), int(H), int(W))
        self.bin, self.n_max = int(bin), int(n_max)
        self.has_illum = illum is not None
        self.label_dtype = np.dtype(label_dtype)
        if self.label_dtype not in (np.dtype(np.int32), np.dtype(np.uint16)):
            raise TypeError("label masks must be int32 or uint16")
        if illum is not None:
            illum = np.ascontiguousarray(illum, dtype=np.float32)
            if illum.shape != (C_, H, W):
                raise ValueError("illum shape %s != %s" % (illum.shape, (C_, H, W)))
        self._h = C.c_void_p()
        capi.call("ips_pipeline_create", C.byref(self._h), fields_per_batch, C_, Z, H, W, bin, n_max,
                  depth, self.label_dtype.itemsize, _hp(illum), float(intensity_scale))

    def output_buffers(self):
        """A dict of page-locked host output arrays for one batch."""
        Fb, Cn, _, H, W = self.shape
        bdt = np.float32 if self.has_illum else np.uint32
        return {
            "maxproj": pinned_empty((Fb, Cn, H, W), np.uint16),
            "binned": pinned_empty((Fb, Cn, H // self.bin, W // self.bin), bdt),
            "n_objects": pinned_empty((Fb,), np.int32),
            "ints": pinned_empty((Fb, self.n_max, 6), np.int32),
            "flts": pinned_empty((Fb, self.n_max, 2 + 5 * Cn), np.float32),
        }

    def _chk(self, a, shape, dtype, name):
        if a is None:
            return
        if a.dtype != np.dtype(dtype) or tuple(a.shape) != tuple(shape) or not a.flags.c_contiguous:
            raise ValueError("%s must be a C-contiguous %s array of shape %s" % (name, np.dtype(dtype), shape))

    def submit(self, raw, labels, out):
        Fb, Cn, Z, H, W = self.shape
        self._chk(raw, (Fb, Cn, Z, H, W), np.uint16, "raw")
        self._chk(labels, (Fb, H, W), self.label_dtype, "labels")
        bdt = np.float32 if self.has_illum else np.uint32
        self._chk(out.get("maxproj"), (Fb, Cn, H, W), np.uint16, "out.maxproj")
        self._chk(out.get("binned"), (Fb, Cn, H // self.bin, W // self.bin), bdt, "out.binned")
        self._chk(out.get("n_objects"), (Fb,), np.int32, "out.n_objects")
        self._chk(out.get("ints"), (Fb, self.n_max, 6), np.int32, "out.ints")
        self._chk(out.get("flts"), (Fb, self.n_max, 2 + 5 * Cn), np.float32, "out.flts")
        return int(capi.call("ips_pipeline_submit", self._h, _hp(raw), _hp(labels), _hp(out.get("maxproj")),
                             _hp(out.get("binned")), _hp(out.get("n_objects")), _hp(out.get("ints")),
                             _hp(out.get("flts"))))

    def wait(self, ticket):
        capi.call("ips_pipeline_wait", self._h, int(ticket))

    def h2d_bytes(self):
        Fb, Cn, Z, H, W = self.shape
        return Fb * (Cn * Z * H * W * 2 + H * W * self.label_dtype.itemsize)

    def d2h_bytes(self, out):
        return int(sum(a.nbytes for a in out.values() if a is not None))

    def close(self):
        if self._h:
            capi.call("ips_pipeline_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
