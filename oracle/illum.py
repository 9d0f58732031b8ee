"""Oracle: per-plate illumination-function estimation (A5).  PARITY UNPINNED.

Test infrastructure only -- see oracle/__init__.py.

The reference only *loads* ``{ch}_illum.npy`` (Illumination_QC_mult.py:186-193); the
functions are produced by CellProfiler pipelines outside the repository.  The definition
restated here is the builder-defined one of SURVEY.md section 8a row A5
(CellProfiler CorrectIlluminationCalculate style): per-pixel mean (or median) over the
plate's max-projected fields, edge-normalised Gaussian smoothing, robust-minimum rescale
so that the function is >= 1.
"""
import numpy as np
import scipy.ndimage as ndi

TRUNCATE = 4.0


def sigma_from_filter_size(filter_size):
    """CellProfiler's FWHM convention: sigma = filter_size / 2.35."""
    return float(filter_size) / 2.35


def accumulate(fields):
    """Exact integer sum over fields: fields[F][C][H][W] uint16 -> uint64 [C][H][W]."""
    return np.asarray(fields).astype(np.uint64).sum(axis=0)


def smooth(raw, sigma):
    """Gaussian smoothing with zero padding, renormalised by the smoothed all-ones mask."""
    raw = np.asarray(raw, dtype=np.float64)
    num = ndi.gaussian_filter(raw, sigma, mode="constant", cval=0.0, truncate=TRUNCATE)
    den = ndi.gaussian_filter(np.ones_like(raw), sigma, mode="constant", cval=0.0,
                              truncate=TRUNCATE)
    return num / den


def robust_rescale(sm, robust_frac=0.02):
    """Divide by the robust minimum (value at the ``robust_frac`` quantile position of
    the sorted positive values) and clamp below at 1."""
    pos = np.sort(sm[sm > 0], axis=None)
    if pos.size == 0:
        return np.ones_like(sm)
    m = pos[int(pos.size * robust_frac)]
    return np.maximum(sm, m) / m


def estimate_from_sum(acc, n_fields, sigma, robust_frac=0.02):
    """acc[C][H][W] integer sums over n_fields -> illum[C][H][W] float64."""
    acc = np.asarray(acc)
    out = np.empty(acc.shape, np.float64)
    for c in range(acc.shape[0]):
        raw = acc[c].astype(np.float64) / float(n_fields)
        out[c] = robust_rescale(smooth(raw, sigma), robust_frac)
    return out


def estimate(fields, sigma, robust_frac=0.02, mode="mean"):
    """fields[F][C][H][W] uint16 -> illum[C][H][W] float64 (mean or per-pixel median)."""
    fields = np.asarray(fields)
    if mode == "mean":
        return estimate_from_sum(accumulate(fields), fields.shape[0], sigma, robust_frac)
    if mode != "median":
        raise ValueError(mode)
    raw = np.median(fields, axis=0)
    out = np.empty(raw.shape, np.float64)
    for c in range(raw.shape[0]):
        out[c] = robust_rescale(smooth(raw[c], sigma), robust_frac)
    return out
