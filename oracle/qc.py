"""Oracle: image-quality metrics of Illumination_QC_mult.py (A6, A7).

Test infrastructure only -- see oracle/__init__.py.  These are restatements; the
reference functions themselves are imported by oracle/make_golden.py and the two are
compared in tests/test_oracle_golden.py.
"""
import numpy as np
import scipy.fft
import scipy.ndimage
import scipy.stats


def percent_maximal(image):
    """100 * (#pixels equal to the image's own maximum) / #pixels.

    Illumination_QC_mult.py:73-95 (no-mask branch; empty image -> 0.0).
    """
    image = np.asarray(image)
    n = image.size
    if n == 0:
        return 0.0
    top = image.max()
    return 100.0 * float(np.count_nonzero(image == top)) / float(n)


def ring_index(h, w):
    """Integer ring label of every FFT bin, Illumination_QC_mult.py:39-43 and :61.

    The squared radius is folded with flips, i.e. row distance is min(i, H-1-i) (not
    min(i, H-i)); the ring label is floor(sqrt(r2)) + 1.
    """
    i = np.arange(h, dtype=np.int64)
    j = np.arange(w, dtype=np.int64)
    di = np.minimum(i, h - 1 - i)
    dj = np.minimum(j, w - 1 - j)
    r2 = di[:, None] ** 2 + dj[None, :] ** 2
    return np.floor(np.sqrt(r2)).astype(np.int64) + 1


def ring_labels(h, w):
    """Ring labels that are summed: 2 .. floor(min(H,W)/8)-1 (:48, :62)."""
    return np.arange(2, int(np.floor(min(h, w) / 8.0)), dtype=np.int64)


def radial_power_spectrum(img):
    """(labels, magnitude ring sums, power ring sums), Illumination_QC_mult.py:31-70.

    Returns Python lists ``[2], [0], [0]`` when there is no ring to sum (min(H,W) < 24),
    exactly as the reference does at :70 -- the caller's ``powersum > 0`` then raises and
    the slope becomes NaN.
    """
    img = np.asarray(img, dtype=np.float64)
    assert img.ndim == 2
    h, w = img.shape
    if np.ptp(img) > 0:                                    # :52-53 MAD normalisation
        img = img / np.median(np.abs(img - img.mean()))
    spec = np.abs(scipy.fft.fft2(img - img.mean()))         # :57 (DC removed, unshifted)
    power = spec ** 2
    rings = ring_index(h, w)
    labels = ring_labels(h, w)
    if labels.size == 0:
        return [2], [0], [0]
    magsum = scipy.ndimage.sum(spec, rings, labels)
    powsum = scipy.ndimage.sum(power, rings, labels)
    return labels, np.asarray(magsum), np.asarray(powsum)


def slope_from_rings(labels, powersum):
    """Least-squares slope of log(power) against log(ring), :108-114."""
    labels = np.asarray(labels)
    powersum = np.asarray(powersum)
    ok = powersum > 0
    if np.count_nonzero(ok) > 2:
        return float(scipy.stats.linregress(np.log(labels[ok]), np.log(powersum[ok]))[0])
    return 0.0


def power_loglog_slope(img):
    """ImageQuality_PowerLogLogSlope, Illumination_QC_mult.py:104-116 (NaN on error)."""
    try:
        labels, _, powersum = radial_power_spectrum(img)
        ok = powersum > 0                                   # TypeError for the list case
        if np.sum(ok) > 2:
            return float(scipy.stats.linregress(np.log(labels[ok]), np.log(powersum[ok]))[0])
        return 0.0
    except Exception:
        return float("nan")


def qc_metrics(img, channel):
    """The two ImageQuality_* entries of one channel, :98-125."""
    return {
        f"ImageQuality_PowerLogLogSlope_{channel}": power_loglog_slope(img),
        f"ImageQuality_PercentMaximal_{channel}": percent_maximal(img),
    }
