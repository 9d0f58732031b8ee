"""Tensor-level entry points: torch tensors in, torch tensors out, arithmetic in libips.so.

PyTorch is used for device memory and streams only; every function below validates its
arguments, allocates outputs / workspace with torch and calls one C-ABI entry point on
the current CUDA stream.  Nothing here computes on the CPU and nothing imports oracle/.
"""
import ctypes as C

import torch

from . import capi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check(t, name, dtype, ndim=None, dev=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s (got %s)" % (name, dtype, t.dtype))
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (there is no CPU path)" % name)
    if dev is not None and t.device != dev:
        raise ValueError("%s is on %s, expected %s" % (name, t.device, dev))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dimensions (got %s)" % (name, ndim, tuple(t.shape)))
    return t


def _workspace(nbytes, dev):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


def device_info():
    sm, ma, mi, l2 = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    capi.call("ips_device_info", C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "l2_bytes": l2.value}


# ---- K1 ---------------------------------------------------------------------------------
def preprocess_fused(raw, illum=None, bin=1, want_maxproj=True, want_corrected=False,
                     want_binned=True, want_pct_maximal=False, out=None):
    """Fused z-max -> illumination divide -> bin x bin sum (K1).

    raw [F][C][Z][H][W] uint16 (device); illum [C][H][W] float32 or None.
    Returns a dict with the requested outputs: ``maxproj`` uint16 [F][C][H][W],
    ``corrected`` float32 [F][C][H][W], ``binned`` [F][C][H/bin][W/bin] (uint32 sums
    stored as int32-viewable ``torch.uint32`` when illum is None, float32 otherwise),
    ``pct_maximal`` float64 [F][C].  ``out`` may carry preallocated tensors under the
    same keys (reused by the benchmark ring).
    Replaces MaxProjection.py:45, Illumination_QC_mult.py:145-150 and :73-95.
    """
    _check(raw, "raw", torch.uint16, 5)
    dev = raw.device
    F, Cn, Z, H, W = raw.shape
    if illum is not None:
        _check(illum, "illum", torch.float32, 3, dev)
        if tuple(illum.shape) != (Cn, H, W):
            raise ValueError("illum shape %s does not match the field (%d, %d, %d)" %
                             (tuple(illum.shape), Cn, H, W))
    if bin not in (1, 2, 4):
        raise ValueError("bin must be 1, 2 or 4")
    if H % bin or W % bin:
        raise ValueError("image size %dx%d is not divisible by bin %d" % (H, W, bin))
    if want_corrected and illum is None:
        raise ValueError("corrected output requires an illumination function")
    out = dict(out or {})
    res = {}
    with torch.cuda.device(dev):
        if want_maxproj:
            res["maxproj"] = out.get("maxproj")
            if res["maxproj"] is None:
                res["maxproj"] = torch.empty((F, Cn, H, W), dtype=torch.uint16, device=dev)
            _check(res["maxproj"], "out.maxproj", torch.uint16, 4, dev)
        if want_corrected:
            res["corrected"] = out.get("corrected")
            if res["corrected"] is None:
                res["corrected"] = torch.empty((F, Cn, H, W), dtype=torch.float32, device=dev)
            _check(res["corrected"], "out.corrected", torch.float32, 4, dev)
        if want_binned:
            bdt = torch.float32 if illum is not None else torch.uint32
            res["binned"] = out.get("binned")
            if res["binned"] is None:
                res["binned"] = torch.empty((F, Cn, H // bin, W // bin), dtype=bdt, device=dev)
            _check(res["binned"], "out.binned", bdt, 4, dev)
        ws, ws_bytes = None, 0
        if want_pct_maximal:
            res["pct_maximal"] = out.get("pct_maximal")
            if res["pct_maximal"] is None:
                res["pct_maximal"] = torch.empty((F, Cn), dtype=torch.float64, device=dev)
            ws_bytes = capi.call("ips_preprocess_workspace_bytes", F, Cn, H, W, bin)
            ws = out.get("ws")
            if ws is None or ws.numel() < ws_bytes:
                ws = _workspace(ws_bytes, dev)
            res["ws"] = ws
        if F > 0:
            capi.call("ips_preprocess_fused", _ptr(raw), _ptr(illum), _ptr(res.get("maxproj")),
                      _ptr(res.get("corrected")), _ptr(res.get("binned")), bin,
                      _ptr(res.get("pct_maximal")), _ptr(ws), ws_bytes, F, Cn, Z, H, W, _stream(dev))
    return res


# ---- K3 ---------------------------------------------------------------------------------
def object_stats(labels, maxproj, illum=None, intensity_scale=1.0, n_max=None, out=None):
    """Per-object statistics over label masks (K3).

    labels [F][H][W] int32, maxproj [F][C][H][W] uint16, illum [C][H][W] float32 or None.
    Returns dict: ``n_objects`` int32 [F], ``ints`` int32 [F][Nmax][6] (label, area, y0,
    x0, y1, x1), ``flts`` float32 [F][Nmax][2+5C] (cy, cx, then per channel sum, mean,
    std, min, max).  Rows beyond n_objects[f] are unspecified.  ``n_max`` bounds the label
    values (default: labels.max(), which costs a device sync).
    Replaces the CellProfiler subprocess of Feature_extraction_opt.py:166-167.
    """
    _check(labels, "labels", torch.int32, 3)
    dev = labels.device
    F, H, W = labels.shape
    _check(maxproj, "maxproj", torch.uint16, 4, dev)
    if maxproj.shape[0] != F or tuple(maxproj.shape[2:]) != (H, W):
        raise ValueError("maxproj shape %s does not match labels %s" % (tuple(maxproj.shape), tuple(labels.shape)))
    Cn = maxproj.shape[1]
    if illum is not None:
        _check(illum, "illum", torch.float32, 3, dev)
        if tuple(illum.shape) != (Cn, H, W):
            raise ValueError("illum shape mismatch")
    if n_max is None:
        n_max = int(labels.max().item()) if labels.numel() else 0
    n_max = max(int(n_max), 1)
    out = dict(out or {})
    with torch.cuda.device(dev):
        n_obj = out.get("n_objects")
        if n_obj is None:
            n_obj = torch.empty((F,), dtype=torch.int32, device=dev)
        ints = out.get("ints")
        if ints is None:
            ints = torch.empty((F, n_max, 6), dtype=torch.int32, device=dev)
        flts = out.get("flts")
        if flts is None:
            flts = torch.empty((F, n_max, 2 + 5 * Cn), dtype=torch.float32, device=dev)
        ws_bytes = capi.call("ips_object_stats_workspace_bytes", F, Cn, n_max)
        ws = out.get("ws")
        if ws is None or ws.numel() < ws_bytes:
            ws = _workspace(ws_bytes, dev)
        if F > 0:
            capi.call("ips_object_stats", _ptr(labels), _ptr(maxproj), _ptr(illum), float(intensity_scale),
                      _ptr(n_obj), _ptr(ints), _ptr(flts), n_max, _ptr(ws), ws.numel(), F, Cn, H, W,
                      _stream(dev))
    return {"n_objects": n_obj, "ints": ints, "flts": flts, "ws": ws}


# ---- K2 ---------------------------------------------------------------------------------
class IllumEstimator:
    """Streaming per-plate illumination-function estimation (K2).

    ``add(maxproj [F][C][H][W] uint16)`` folds a batch of max-projected fields into exact
    uint32 per-pixel sums; ``finalize(sigma, robust_frac)`` returns the function
    [C][H][W] float32 (mean -> edge-normalised Gaussian -> robust-minimum rescale, >= 1).
    Produces what Illumination_QC_mult.py:186-193 loads as ``{ch}_illum.npy``.
    """

    MAX_FIELDS = 65537      # 65537 * 65535 < 2^32: the uint32 sums cannot overflow

    def __init__(self, C_, H, W, device="cuda"):
        self.dev = torch.device(device)
        self.shape = (int(C_), int(H), int(W))
        self.acc = torch.zeros(self.shape, dtype=torch.uint32, device=self.dev)
        self.n = 0

    def add(self, maxproj):
        _check(maxproj, "maxproj", torch.uint16, 4, self.acc.device)
        if tuple(maxproj.shape[1:]) != self.shape:
            raise ValueError("maxproj shape %s does not match %s" % (tuple(maxproj.shape), self.shape))
        F = maxproj.shape[0]
        if self.n + F > self.MAX_FIELDS:
            raise OverflowError("more than %d fields would overflow the uint32 sums" % self.MAX_FIELDS)
        if F:
            with torch.cuda.device(self.acc.device):
                capi.call("ips_illum_accumulate", _ptr(maxproj), _ptr(self.acc), F, *self.shape,
                          _stream(self.acc.device))
        self.n += F
        return self

    def finalize(self, sigma, robust_frac=0.02):
        if self.n == 0:
            raise ValueError("no fields accumulated")
        Cn, H, W = self.shape
        dev = self.acc.device
        with torch.cuda.device(dev):
            out = torch.empty(self.shape, dtype=torch.float32, device=dev)
            nbytes = capi.call("ips_illum_finalize_workspace_bytes", Cn, H, W)
            ws = _workspace(nbytes, dev)
            capi.call("ips_illum_finalize", _ptr(self.acc), self.n, float(sigma), float(robust_frac),
                      _ptr(out), _ptr(ws), ws.numel(), Cn, H, W, _stream(dev))
        return out


def illum_median(stack):
    """Per-pixel median over a device-resident stack [N][C][H][W] uint16 -> float32 [C][H][W]."""
    _check(stack, "stack", torch.uint16, 4)
    N, Cn, H, W = stack.shape
    if N == 0:
        raise ValueError("empty stack")
    dev = stack.device
    with torch.cuda.device(dev):
        out = torch.empty((Cn, H, W), dtype=torch.float32, device=dev)
        capi.call("ips_illum_median", _ptr(stack), _ptr(out), N, Cn, H, W, _stream(dev))
    return out


def illum_smooth_rescale(raw, sigma, robust_frac=0.02):
    """Gaussian smoothing + robust-minimum rescale of a raw function [C][H][W] float32."""
    _check(raw, "raw", torch.float32, 3)
    Cn, H, W = raw.shape
    dev = raw.device
    with torch.cuda.device(dev):
        out = torch.empty_like(raw)
        ws = _workspace(capi.call("ips_illum_finalize_workspace_bytes", Cn, H, W), dev)
        capi.call("ips_illum_smooth_rescale", _ptr(raw), float(sigma), float(robust_frac), _ptr(out),
                  _ptr(ws), ws.numel(), Cn, H, W, _stream(dev))
    return out


# ---- K5 ---------------------------------------------------------------------------------
def lanczos_resize_u16(planes, out_hw):
    """Pillow-exact LANCZOS resize of uint16 planes [C][H][W] -> [C][outH][outW] (K5).
    Replaces ``img.resize(target, resample=LANCZOS)`` of Image_re-binning.py:18."""
    _check(planes, "planes", torch.uint16, 3)
    Cn, H, W = planes.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    if oh <= 0 or ow <= 0:
        raise ValueError("target size must be positive")
    dev = planes.device
    with torch.cuda.device(dev):
        out = torch.empty((Cn, oh, ow), dtype=torch.uint16, device=dev)
        if Cn:
            ws = _workspace(capi.call("ips_lanczos_workspace_bytes", Cn, H, W, oh, ow), dev)
            capi.call("ips_lanczos_resize_u16", _ptr(planes), _ptr(out), Cn, H, W, oh, ow, _ptr(ws),
                      ws.numel(), _stream(dev))
    return out


# ---- K6 ---------------------------------------------------------------------------------
def ring_sums(spec, n_rings):
    """Ring sums of |z| and |z|^2 over the folded FFT radius (Illumination_QC_mult.py:61-68).
    spec [F][H][W] complex128 -> (mag [F][n_rings], pow [F][n_rings]) float64."""
    _check(spec, "spec", torch.complex128, 3)
    F, H, W = spec.shape
    dev = spec.device
    with torch.cuda.device(dev):
        mag = torch.empty((F, n_rings), dtype=torch.float64, device=dev)
        pw = torch.empty((F, n_rings), dtype=torch.float64, device=dev)
        if F and n_rings > 0:
            capi.call("ips_ring_sums", _ptr(spec), _ptr(mag), _ptr(pw), int(n_rings), F, H, W, _stream(dev))
    return mag, pw


def rps_spectrum(image, illum=None, want_corrected=False):
    """rps() of Illumination_QC_mult.py:31-70 up to the ring sums, on the device: image [H][W]
    uint16 or float64, illum [H][W] float64 or None (the float64 divide of :145-150).
    Returns (mag [n_rings], pow [n_rings]) float64 device tensors for ring labels 2 .. n_rings + 1
    (empty when min(H, W) < 24), plus the float64 corrected image when ``want_corrected``.
    Median by exact radix select, real FFT (cuFFT via torch.fft.rfft2 is the FFT library), ring
    sums over the Hermitian half -- no host round trip."""
    if not isinstance(image, torch.Tensor) or image.dtype not in (torch.uint16, torch.float64):
        raise TypeError("image must be a uint16 or float64 tensor")
    _check(image, "image", image.dtype, 2)
    dev = image.device
    H, W = image.shape
    if illum is not None:
        _check(illum, "illum", torch.float64, 2, dev)
        if tuple(illum.shape) != (H, W):
            raise ValueError("illum shape mismatch")
    import math
    n_rings = max(int(math.floor(min(H, W) / 8.0)) - 2, 0)
    with torch.cuda.device(dev):
        z = torch.empty((H, W), dtype=torch.float64, device=dev)
        computed = image.dtype == torch.uint16 or illum is not None
        corrected = torch.empty((H, W), dtype=torch.float64, device=dev) if (computed and want_corrected) else (z if computed else None)
        ws = _workspace(capi.call("ips_rps_prepare_workspace_bytes", H * W), dev)
        capi.call("ips_rps_prepare", _ptr(image) if image.dtype == torch.uint16 else None,
                  _ptr(image) if image.dtype == torch.float64 else None, _ptr(illum), _ptr(corrected), _ptr(z), H * W,
                  _ptr(ws), ws.numel(), _stream(dev))
        mag = torch.empty((n_rings,), dtype=torch.float64, device=dev)
        pw = torch.empty((n_rings,), dtype=torch.float64, device=dev)
        if n_rings > 0:
            spec = torch.fft.rfft2(z).contiguous()         # cuFFT's two-pass result can come back strided
            capi.call("ips_ring_sums_half", _ptr(spec), _ptr(mag), _ptr(pw), n_rings, 1, H, W, _stream(dev))
    if want_corrected:
        return mag, pw, (corrected if computed else image)
    return mag, pw


def loglog_slope(powersum):
    """Least-squares slope of log(power) over log(ring label) (labels 2..), rings with power > 0,
    0.0 with fewer than three (Illumination_QC_mult.py:108-114).  powersum [F][n_rings] or
    [n_rings] float64 (device) -> float64 device tensor [F]."""
    _check(powersum, "powersum", torch.float64)
    p2 = powersum.reshape(-1, powersum.shape[-1])
    dev = powersum.device
    F, n = p2.shape
    with torch.cuda.device(dev):
        out = torch.zeros((F,), dtype=torch.float64, device=dev)
        if n > 0 and F > 0:
            capi.call("ips_loglog_slope", _ptr(p2.contiguous()), _ptr(out), n, F, _stream(dev))
    return out


# ---- well aggregation -------------------------------------------------------------------
def well_mean(rows, well, n_wells):
    """Per-well mean of object rows [N][D] float32 with well ids [N] int32 in [0, n_wells).
    Returns (mean [n_wells][D] float64, NaN for empty wells; count [n_wells] int32).
    Replaces groupby('Metadata_Well').agg('mean') of Normalize_CP_ami.py:126."""
    _check(rows, "rows", torch.float32, 2)
    dev = rows.device
    _check(well, "well", torch.int32, 1, dev)
    N, D = rows.shape
    if well.shape[0] != N:
        raise ValueError("one well id per row expected")
    with torch.cuda.device(dev):
        mean = torch.empty((n_wells, D), dtype=torch.float64, device=dev)
        count = torch.empty((n_wells,), dtype=torch.int32, device=dev)
        ws = _workspace(capi.call("ips_well_mean_workspace_bytes", n_wells, D), dev)
        capi.call("ips_well_mean", _ptr(rows), _ptr(well), _ptr(mean), _ptr(count), N, D, n_wells,
                  _ptr(ws), ws.numel(), _stream(dev))
    return mean, count


def well_mean_f64(rows, well, n_wells):
    """As ``well_mean`` on float64 rows (the CellProfiler tables the scripts read)."""
    _check(rows, "rows", torch.float64, 2)
    dev = rows.device
    _check(well, "well", torch.int32, 1, dev)
    N, D = rows.shape
    if well.shape[0] != N:
        raise ValueError("one well id per row expected")
    with torch.cuda.device(dev):
        mean = torch.empty((n_wells, D), dtype=torch.float64, device=dev)
        count = torch.empty((n_wells,), dtype=torch.int32, device=dev)
        ws = _workspace(capi.call("ips_well_mean_workspace_bytes", n_wells, D), dev)
        capi.call("ips_well_mean_f64", _ptr(rows), _ptr(well), _ptr(mean), _ptr(count), N, D, n_wells,
                  _ptr(ws), ws.numel(), _stream(dev))
    return mean, count


def well_median_f64(rows, well, n_wells):
    """Per-well median of float64 rows [N][D] (NaN skipped; mean of the two middle values):
    ``groupby('Metadata_Well').agg('median')``, Normalize_CP_ami.py:126 with
    ``--well_agg_func median`` (:163).  Returns (median [n_wells][D] float64, count [n_wells])."""
    _check(rows, "rows", torch.float64, 2)
    dev = rows.device
    _check(well, "well", torch.int32, 1, dev)
    N, D = rows.shape
    if well.shape[0] != N:
        raise ValueError("one well id per row expected")
    with torch.cuda.device(dev):
        # grouping the row indices by well is index plumbing (torch.sort); the selection is ours
        w64 = well.to(torch.int64)
        inside = (w64 >= 0) & (w64 < n_wells)
        key = torch.where(inside, w64, torch.full_like(w64, n_wells))
        perm = torch.sort(key, stable=True).indices.contiguous()
        counts = torch.bincount(key, minlength=n_wells + 1)[:n_wells]
        offsets = torch.zeros((n_wells + 1,), dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(counts, 0)
        med = torch.empty((n_wells, D), dtype=torch.float64, device=dev)
        count = torch.empty((n_wells,), dtype=torch.int32, device=dev)
        capi.call("ips_well_median_f64", _ptr(rows), _ptr(perm), _ptr(offsets), _ptr(med), _ptr(count), N, D,
                  n_wells, _stream(dev))
    return med, count


# ---- K1 + K3 in one pass ------------------------------------------------------------------
def illum_reciprocal(illum):
    """1 / illum (IEEE division, once per plate): pass the result as ``illum_rcp`` to
    ``field_fused`` and the per-pixel divide becomes a multiply."""
    _check(illum, "illum", torch.float32)
    dev = illum.device
    with torch.cuda.device(dev):
        out = torch.empty_like(illum)
        capi.call("ips_illum_reciprocal", _ptr(illum), _ptr(out), illum.numel(), _stream(dev))
    return out


def field_fused(raw, illum, labels, bin=2, intensity_scale=1.0, n_max=None, want_maxproj=True,
                want_binned=True, out=None, illum_rcp=None):
    """One pass over a batch of fields: max projection, b x b sum bin (of the corrected image
    when illum is given) and per-object statistics over the label masks.

    raw [F][C][Z][H][W] uint16, illum [C][H][W] float32 or None, labels [F][H][W] int32 or
    uint16 (Cellpose's own dtype below 65536 objects).  ``illum_rcp`` = ``illum_reciprocal(illum)``
    may be given instead of / next to ``illum`` (computed once per plate).
    Returns dict with ``maxproj``, ``binned`` (as preprocess_fused) and ``n_objects``,
    ``ints``, ``flts`` (as object_stats).  Same results as the two separate calls.
    """
    _check(raw, "raw", torch.uint16, 5)
    dev = raw.device
    F, Cn, Z, H, W = raw.shape
    if not isinstance(labels, torch.Tensor) or labels.dtype not in (torch.int32, torch.uint16):
        raise TypeError("labels must be an int32 or uint16 tensor")
    _check(labels, "labels", labels.dtype, 3, dev)
    if tuple(labels.shape) != (F, H, W):
        raise ValueError("labels shape %s does not match raw %s" % (tuple(labels.shape), tuple(raw.shape)))
    func, is_rcp = illum, 0
    if illum_rcp is not None:
        func, is_rcp = illum_rcp, 1
    if func is not None:
        _check(func, "illum", torch.float32, 3, dev)
        if tuple(func.shape) != (Cn, H, W):
            raise ValueError("illum shape mismatch")
    if bin not in (1, 2, 4):
        raise ValueError("bin must be 1, 2 or 4")
    if H % bin or W % bin:
        raise ValueError("image size %dx%d is not divisible by bin %d" % (H, W, bin))
    if n_max is None:
        n_max = int(labels.to(torch.int32).max().item()) if labels.numel() else 0
    n_max = max(int(n_max), 1)
    out = dict(out or {})
    with torch.cuda.device(dev):
        mp = out.get("maxproj")
        if mp is None and (want_maxproj or W % 8):
            mp = torch.empty((F, Cn, H, W), dtype=torch.uint16, device=dev)
        bn = out.get("binned")
        if bn is None and want_binned:
            bn = torch.empty((F, Cn, H // bin, W // bin),
                             dtype=torch.float32 if func is not None else torch.uint32, device=dev)
        n_obj = out.get("n_objects")
        if n_obj is None:
            n_obj = torch.empty((F,), dtype=torch.int32, device=dev)
        ints = out.get("ints")
        if ints is None:
            ints = torch.empty((F, n_max, 6), dtype=torch.int32, device=dev)
        flts = out.get("flts")
        if flts is None:
            flts = torch.empty((F, n_max, 2 + 5 * Cn), dtype=torch.float32, device=dev)
        ws_bytes = capi.call("ips_field_fused_workspace_bytes", F, Cn, H, W, bin, n_max)
        ws = out.get("ws")
        if ws is None or ws.numel() < ws_bytes:
            ws = _workspace(ws_bytes, dev)
        if F > 0:
            capi.call("ips_field_fused_ex", _ptr(raw), _ptr(func), is_rcp, _ptr(labels), labels.element_size(),
                      _ptr(mp), _ptr(bn), bin, float(intensity_scale), _ptr(n_obj), _ptr(ints), _ptr(flts), n_max,
                      _ptr(ws), ws.numel(), F, Cn, Z, H, W, _stream(dev))
    return {"maxproj": mp, "binned": bn, "n_objects": n_obj, "ints": ints, "flts": flts, "ws": ws}


# ---- K4 ---------------------------------------------------------------------------------
def cosine_triu(X, group=None, n_groups=None):
    """Per-group sum of strict-upper-triangle cosine similarities (K4).

    X [N][D] float32 (NaN already replaced by 0); group [N] int32 with ascending ids
    0..n_groups-1 (rows of a group contiguous) or None for a single group.
    Returns (sum [n_groups] float64, npairs [n_groups] int64); mean = sum / npairs.
    Replaces cosine_similarity -> triu(k=1) of Feature_select_cosine_ami.py:145-149.
    """
    _check(X, "X", torch.float32, 2)
    dev = X.device
    N, D = X.shape
    if group is not None:
        _check(group, "group", torch.int32, 1, dev)
        if group.shape[0] != N:
            raise ValueError("one group id per row expected")
        if n_groups is None:
            n_groups = int(group.max().item()) + 1 if N else 1
    else:
        n_groups = 1
    with torch.cuda.device(dev):
        s = torch.empty((n_groups,), dtype=torch.float64, device=dev)
        npairs = torch.empty((n_groups,), dtype=torch.int64, device=dev)
        ws = _workspace(capi.call("ips_cosine_workspace_bytes", max(N, 1), D), dev)
        capi.call("ips_cosine_triu", _ptr(X), _ptr(group), int(n_groups), _ptr(s), _ptr(npairs), N, D,
                  _ptr(ws), ws.numel(), _stream(dev))
    return s, npairs


# ---- profile normalisation -----------------------------------------------------------------
def mad_robustize(profiles, is_control):
    """(x - median_ctrl) / (1.4826 * MAD_ctrl + 1e-18) per feature column (device, float64).
    profiles [W][D] float64, is_control [W] bool/uint8.  Replaces pycytominer's
    normalize(method="mad_robustize") at Normalize_CP_ami.py:137-142."""
    _check(profiles, "profiles", torch.float64, 2)
    dev = profiles.device
    ctrl = is_control.to(device=dev, dtype=torch.uint8).contiguous()
    W, D = profiles.shape
    if ctrl.shape[0] != W:
        raise ValueError("one control flag per well expected")
    with torch.cuda.device(dev):
        out = torch.empty_like(profiles)
        capi.call("ips_mad_robustize", _ptr(profiles), _ptr(ctrl), _ptr(out), W, D, _stream(dev))
    return out


def double_sigmoid_abs(x, k=3, alpha=2.3538):
    """|double_sigmoid(x)| elementwise (Feature_select_cosine_ami.py:22-27, :117-118)."""
    _check(x, "x", torch.float64)
    dev = x.device
    with torch.cuda.device(dev):
        y = torch.empty_like(x)
        capi.call("ips_double_sigmoid_abs", _ptr(x), _ptr(y), x.numel(), int(k), float(alpha), _stream(dev))
    return y


# ---- cell crops ----------------------------------------------------------------------------
def cell_crops(corrected, labels, ints, n_objects, box=200, max_crops=None):
    """Centroid-centred masked crops of every object, min-max scaled to uint8 per crop and
    channel (Cellpose_GPU_s3fs.py:149-182, :34-43).

    corrected [F][C][H][W] float32, labels [F][H][W] int32, ints [F][Nmax][6] / n_objects [F]
    from object_stats / field_fused.  Returns dict: ``crops`` uint8 [F][max_crops][C][box][box],
    ``kept`` int32 [F][max_crops][3] (label, yc, xc), ``n_kept`` int32 [F].
    """
    _check(corrected, "corrected", torch.float32, 4)
    dev = corrected.device
    F, Cn, H, W = corrected.shape
    _check(labels, "labels", torch.int32, 3, dev)
    _check(ints, "ints", torch.int32, 3, dev)
    _check(n_objects, "n_objects", torch.int32, 1, dev)
    if tuple(labels.shape) != (F, H, W) or ints.shape[0] != F or ints.shape[2] != 6 or n_objects.shape[0] != F:
        raise ValueError("shape mismatch between corrected, labels and object rows")
    n_max = ints.shape[1]
    if max_crops is None:
        max_crops = n_max
    with torch.cuda.device(dev):
        crops = torch.empty((F, max_crops, Cn, box, box), dtype=torch.uint8, device=dev)
        kept = torch.zeros((F, max_crops, 3), dtype=torch.int32, device=dev)
        n_kept = torch.empty((F,), dtype=torch.int32, device=dev)
        ws = _workspace(capi.call("ips_cell_crops_workspace_bytes", F, n_max), dev)
        capi.call("ips_cell_crops", _ptr(corrected), _ptr(labels), _ptr(ints), _ptr(n_objects), int(box),
                  int(max_crops), _ptr(crops), _ptr(n_kept), _ptr(kept), _ptr(ws), ws.numel(), n_max, F, Cn, H, W,
                  _stream(dev))
    return {"crops": crops, "kept": kept, "n_kept": n_kept}


def cosine_triu_pairs(X, group=None, group_sizes=None):
    """Per-group strict-upper-triangle similarities themselves (Pycyto_pertime.py:150-155).

    X [N][D] float32, group [N] int32 ascending contiguous ids (or None), ``group_sizes`` the
    host-side list of group sizes (needed to size the output; default: one group of N rows).
    Returns (sum [G] float64, npairs [G] int64, pairs float64 [total], offsets int64 [G + 1]).
    """
    _check(X, "X", torch.float32, 2)
    dev = X.device
    N, D = X.shape
    if group is None:
        group_sizes = [N]
    elif group_sizes is None:
        group_sizes = torch.bincount(group.to(torch.int64)).cpu().tolist()
    G = len(group_sizes)
    total = int(sum(n * (n - 1) // 2 for n in group_sizes))
    if group is not None:
        _check(group, "group", torch.int32, 1, dev)
    with torch.cuda.device(dev):
        s = torch.empty((G,), dtype=torch.float64, device=dev)
        npairs = torch.empty((G,), dtype=torch.int64, device=dev)
        pairs = torch.empty((max(total, 1),), dtype=torch.float64, device=dev)
        offsets = torch.empty((G + 1,), dtype=torch.int64, device=dev)
        ws = _workspace(capi.call("ips_cosine_pairs_workspace_bytes", N, D), dev)
        capi.call("ips_cosine_triu_pairs", _ptr(X), _ptr(group), G, _ptr(s), _ptr(npairs), _ptr(pairs), _ptr(offsets),
                  max(total, N * (N - 1) // 2 if G == 1 else total), N, D, _ptr(ws), ws.numel(), _stream(dev))
    return s, npairs, pairs[:total], offsets


# ---- K7: TIFF strip codec -----------------------------------------------------------------
def tiff_rows_per_strip(H, W):
    """Pillow's strip height for an H x W 16-bit plane."""
    return int(capi.call("ips_tiff_rows_per_strip", int(H), int(W)))


def tiff_lzw_encode(planes, rows_per_strip=None):
    """uint16 planes [P][H][W] (device) -> (files uint8 [P][file_cap] on the device, file_bytes
    int64 [P] on the host).  files[p, :file_bytes[p]] is the complete LZW TIFF of plane p, the
    bytes ``img.save(buf, format='TIFF', compression='tiff_lzw')`` (Image_re-binning.py:19-21)
    produces for the same pixels."""
    _check(planes, "planes", torch.uint16, 3)
    P, H, W = planes.shape
    if H <= 0 or W <= 0:
        raise ValueError("empty planes")
    dev = planes.device
    rps = tiff_rows_per_strip(H, W) if rows_per_strip is None else int(rows_per_strip)
    if rps <= 0:
        raise ValueError("rows_per_strip must be positive")
    cap = int(capi.call("ips_tiff_file_bound", H, W, rps))
    with torch.cuda.device(dev):
        files = torch.empty((P, cap), dtype=torch.uint8, device=dev)
        nbytes = torch.zeros((P,), dtype=torch.int64, device=dev)
        if P:
            ws = _workspace(capi.call("ips_tiff_encode_workspace_bytes", P, H, W, rps), dev)
            capi.call("ips_tiff_lzw_encode_u16", _ptr(planes), P, H, W, rps, _ptr(files), cap, _ptr(nbytes),
                      _ptr(ws), ws.numel(), _stream(dev))
        host = nbytes.cpu()
    if P and int(host.min()) <= 0:
        raise capi.IpsError(-6, "ips_tiff_lzw_encode_u16: a file did not fit its slot")
    return files, host


def tiff_lzw_decode(src, src_off, src_bytes, dst, dst_off, dst_bytes, to_host=True):
    """LZW strips src[src_off[s] : +src_bytes[s]] -> dst[dst_off[s] : +dst_bytes[s]] (uint8 views,
    device).  The descriptor arrays are host sequences or NumPy arrays.  Returns the per-strip status (host
    int32 tensor; 0 = ok).  Replaces the strip decoding behind Image.open / imageio.imread
    (Image_re-binning.py:17, MaxProjection.py:39)."""
    _check(src, "src", torch.uint8, 1)
    _check(dst, "dst", torch.uint8, 1, src.device)
    n = len(src_off)
    if not (len(src_bytes) == len(dst_off) == len(dst_bytes) == n):
        raise ValueError("descriptor arrays differ in length")
    if n == 0:
        return torch.zeros((0,), dtype=torch.int32)
    import numpy as np
    so, sb, do, db = (torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64)) for x in (src_off, src_bytes, dst_off, dst_bytes))
    if int(so.min()) < 0 or int(sb.min()) < 0 or int((so + sb).max()) > src.numel():
        raise ValueError("a source strip lies outside src")
    if int(do.min()) < 0 or int(db.min()) < 0 or int((do + db).max()) > dst.numel():
        raise ValueError("a destination strip lies outside dst")
    if int(sb.max()) >= 2 ** 32 or int(db.max()) >= 2 ** 32:
        raise ValueError("strips of 4 GiB or more are not supported")
    dev = src.device
    with torch.cuda.device(dev):
        # page-locked staging + asynchronous copies: a pageable .to(dev) synchronises the stream, i.e. waits
        # for the whole batch of file bytes queued in front of it (a pipelined plate loop must not block here)
        desc64 = torch.stack([so, do]).pin_memory().to(dev, non_blocking=True)
        desc32 = torch.stack([sb, db]).to(torch.uint32).pin_memory().to(dev, non_blocking=True)
        status = torch.empty((n,), dtype=torch.int32, device=dev)
        capi.call("ips_tiff_lzw_decode", _ptr(src), _ptr(desc64[0]), _ptr(desc32[0]), _ptr(dst), _ptr(desc64[1]),
                  _ptr(desc32[1]), n, _ptr(status), _stream(dev))
        return status.cpu() if to_host else status          # to_host=False: the caller reads it after its own sync


def tiff_fix_u16(img, predictor=1, byteswap=False):
    """In place on uint16 rows [..., W]: byte swap, then undo horizontal differencing (Predictor 2)."""
    _check(img, "img", torch.uint16)
    W = img.shape[-1]
    rows = img.numel() // W if W else 0
    capi.call("ips_tiff_fix_u16", _ptr(img), rows, W, int(predictor), int(bool(byteswap)), _stream(img.device))
    return img
