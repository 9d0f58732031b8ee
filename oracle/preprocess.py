"""Oracle: z-max projection, illumination divide, sum binning (A1, A2, A3').

Test infrastructure only -- see oracle/__init__.py.
"""
import numpy as np


def max_projection(planes):
    """Elementwise maximum over the z planes of one channel of one field.

    Follows MaxProjection.py:42-45: every plane must have the shape of the first
    (ValueError otherwise) and the result is ``np.maximum.reduce`` of the list, which
    keeps the input dtype (uint16 in, uint16 out).
    """
    planes = [np.asarray(p) for p in planes]
    first = planes[0].shape
    for p in planes[1:]:
        if p.shape != first:
            raise ValueError("Image shape mismatch in group")
    out = planes[0].copy()
    for p in planes[1:]:
        np.maximum(out, p, out=out)
    return out


def max_projection_field(raw):
    """raw[C][Z][H][W] -> [C][H][W]; the MaxProjection.py:84-91 loop over channels."""
    raw = np.asarray(raw)
    return raw.max(axis=1)


def illum_correct(img, illum):
    """Illumination correction, Illumination_QC_mult.py:145-153.

    The image is promoted to float64 and divided by the illumination function when one
    is given *and* has the image's shape; a shape mismatch silently leaves the image
    uncorrected (``:151-153``).  ``Cellpose_GPU_s3fs.py:72`` performs the same divide.
    """
    out = np.asarray(img).astype(np.float64)
    if illum is not None and np.shape(illum) == out.shape:
        out = out / np.asarray(illum)
    return out


def sum_bin(img, b):
    """b x b sum binning (SURVEY.md section 8a row A3'; north_star only, no reference site).

    Integer input -> uint32 sums (exact), float input -> float64 sums.
    """
    img = np.asarray(img)
    h, w = img.shape[-2:]
    if h % b or w % b:
        raise ValueError("image size not divisible by bin")
    lead = img.shape[:-2]
    acc = np.uint32 if np.issubdtype(img.dtype, np.integer) else np.float64
    v = img.reshape(lead + (h // b, b, w // b, b)).astype(acc)
    return v.sum(axis=(-3, -1), dtype=acc)


def preprocess_field(raw, illum, b):
    """The fused pass K1 restated on the CPU for one field.

    raw[C][Z][H][W] uint16, illum[C][H][W] float or None.
    Returns (maxproj uint16 [C][H][W], corrected float64 [C][H][W] or None,
             binned [C][H/b][W/b]: uint32 sums of maxproj if illum is None else float64
             sums of corrected).
    """
    mp = max_projection_field(raw)
    if illum is None:
        return mp, None, sum_bin(mp, b)
    corr = mp.astype(np.float64) / np.asarray(illum, dtype=np.float64)
    return mp, corr, sum_bin(corr, b)
