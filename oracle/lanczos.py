"""Oracle: Pillow-LANCZOS "re-binning" of 16-bit images (A3).

Test infrastructure only -- see oracle/__init__.py.

Image_re-binning.py:18 calls ``PIL.Image.resize(target, resample=LANCZOS)`` on the
decoded TIFF (mode ``I;16`` for 16-bit data).  Pillow is installed here, so
``pil_resize`` *is* the reference arithmetic; ``resize_restated`` spells the same
arithmetic out in NumPy (this is what the CUDA kernel implements) and is checked
bit-for-bit against Pillow in tests/test_oracle_golden.py.  Parity is pinned to
Pillow 12.2.0 (the reference's requirements.txt does not pin Pillow).
"""
import math

import numpy as np

LANCZOS_SUPPORT = 3.0


def pil_resize(img_u16, out_hw):
    """The real thing: Pillow's I;16 LANCZOS resize.  out_hw = (out_h, out_w)."""
    from PIL import Image
    im = Image.fromarray(np.ascontiguousarray(img_u16, dtype=np.uint16))
    out = im.resize((int(out_hw[1]), int(out_hw[0])), resample=Image.Resampling.LANCZOS)
    return np.asarray(out, dtype=np.uint16)


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def coefficients(in_size, out_size):
    """Per-output-index window start, length and normalised double weights.

    Pillow's precompute_coeffs: scale = in/out, filterscale = max(scale, 1),
    support = 3 * filterscale, ksize = ceil(support) * 2 + 1,
    center = (xx + 0.5) * scale, xmin = max(int(center - support + 0.5), 0),
    xmax = min(int(center + support + 0.5), in), w_k = L((k + xmin - center + 0.5) / filterscale),
    weights divided by their sum.  Unused tail of each row is zero.
    """
    scale = float(in_size) / float(out_size)
    fscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    xlen = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    inv = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        ww = 0.0
        for k in range(n):
            w = _lanczos((k + lo - center + 0.5) * inv)
            kk[xx, k] = w
            ww += w
        if ww != 0.0:
            for k in range(n):
                kk[xx, k] /= ww
        xmin[xx] = lo
        xlen[xx] = n
    return xmin, xlen, kk


def _store_u16(ss):
    """Pillow's two-byte store: round half away from zero, then clip each byte on its
    own -- negative -> 0, > 65535 -> 0xFF00 | (v & 0xFF)."""
    v = np.where(ss >= 0.0, ss + 0.5, ss - 0.5).astype(np.int64)   # C truncation
    out = np.where(v < 0, 0, v)
    out = np.where(out > 65535, 0xFF00 | (out & 0xFF), out)
    return out.astype(np.uint16)


def _pass_rows(img, xmin, xlen, kk):
    """Resample along the last axis; sequential double accumulation, tap by tap."""
    n_out, ksize = kk.shape
    in_size = img.shape[-1]
    src = img.astype(np.float64)
    acc = np.zeros(img.shape[:-1] + (n_out,), np.float64)
    for t in range(ksize):
        active = t < xlen
        idx = np.minimum(xmin + t, in_size - 1)
        term = src[..., idx] * kk[:, t]
        acc = np.where(active, acc + term, acc)
    return _store_u16(acc)


def resize_restated(img_u16, out_hw):
    """NumPy restatement: horizontal pass, round to uint16, then vertical pass."""
    img = np.ascontiguousarray(img_u16, dtype=np.uint16)
    h, w = img.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    cur = img
    if ow != w:
        cur = _pass_rows(cur, *coefficients(w, ow))
    if oh != h:
        cur = _pass_rows(cur.T, *coefficients(h, oh)).T
    return np.ascontiguousarray(cur)
