"""Oracle: per-object features over a label mask (A4).  PARITY UNPINNED.

Test infrastructure only -- see oracle/__init__.py.

The reference delegates this arithmetic to CellProfiler 4.2.8 inside docker
(Feature_extraction_opt.py:166-167); its source is not under /root/reference and no
reference test touches it.  This restatement uses ``scipy.ndimage`` labelled statistics
-- the primitive CellProfiler's MeasureObjectIntensity / MeasureObjectSizeShape are
built on, and the one the reference itself uses for ring sums at
Illumination_QC_mult.py:66-67 -- with the conventions SURVEY.md section 8c fixes:
population standard deviation (ddof=0), half-open bounding boxes (as
``regionprops.bbox`` used at Cellpose_GPU_s3fs.py:149-166), centroid = mean of integer
pixel (row, col) indices, label 0 = background, rows in ascending label order for the
labels that occur (``regionprops`` skips absent labels, Cellpose_GPU_s3fs.py:149).
"""
import numpy as np
import scipy.ndimage as ndi

INT_COLS = ("label", "area", "y0", "x0", "y1", "x1")


def flt_cols(n_channels):
    cols = ["cy", "cx"]
    for c in range(n_channels):
        cols += [f"sum{c}", f"mean{c}", f"std{c}", f"min{c}", f"max{c}"]
    return tuple(cols)


def intensity_planes(maxproj, illum, intensity_scale):
    """values[C][H][W] float64 = maxproj / illum * scale (divide as in A2)."""
    v = np.asarray(maxproj).astype(np.float64)
    if illum is not None:
        v = v / np.asarray(illum, dtype=np.float64)
    return v * float(intensity_scale)


def object_stats(labels, maxproj, illum=None, intensity_scale=1.0):
    """Per-object rows of one field.

    labels[H][W] integer, maxproj[C][H][W] uint16, illum[C][H][W] or None.
    Returns (ints int64 [N][6] = label, area, y0, x0, y1, x1;
             flts float64 [N][2+5C] = cy, cx, then per channel sum, mean, std, min, max).
    """
    labels = np.asarray(labels)
    vals = intensity_planes(maxproj, illum, intensity_scale)
    C = vals.shape[0]
    counts = np.bincount(labels.ravel().astype(np.int64))
    present = np.flatnonzero(counts[1:] > 0) + 1 if counts.size > 1 else np.zeros(0, np.int64)
    n = present.size
    ints = np.zeros((n, 6), np.int64)
    flts = np.zeros((n, 2 + 5 * C), np.float64)
    if n == 0:
        return ints, flts
    ints[:, 0] = present
    ints[:, 1] = counts[present]
    slices = ndi.find_objects(labels.astype(np.int32), max_label=int(present[-1]))
    for r, lab in enumerate(present):
        sy, sx = slices[lab - 1]
        ints[r, 2:] = (sy.start, sx.start, sy.stop, sx.stop)
    com = np.asarray(ndi.center_of_mass(np.ones(labels.shape), labels, present))
    flts[:, 0:2] = com.reshape(n, 2)
    for c in range(C):
        o = 2 + 5 * c
        flts[:, o + 0] = ndi.sum(vals[c], labels, present)
        flts[:, o + 1] = ndi.mean(vals[c], labels, present)
        flts[:, o + 2] = ndi.standard_deviation(vals[c], labels, present)
        flts[:, o + 3] = ndi.minimum(vals[c], labels, present)
        flts[:, o + 4] = ndi.maximum(vals[c], labels, present)
    return ints, flts


def object_stats_bincount(labels, maxproj, illum=None, intensity_scale=1.0):
    """Independent restatement with sorting / reduceat -- cross-check for the above.

    Same outputs; used by tests to make sure the scipy.ndimage conventions above are the
    ones we think they are, and as a faster checker at full plate sizes.
    """
    labels = np.asarray(labels)
    H, W = labels.shape
    vals = intensity_planes(maxproj, illum, intensity_scale)
    C = vals.shape[0]
    flat = labels.ravel().astype(np.int64)
    fg = np.flatnonzero(flat > 0)
    order = fg[np.argsort(flat[fg], kind="stable")]
    lab_sorted = flat[order]
    if lab_sorted.size == 0:
        return np.zeros((0, 6), np.int64), np.zeros((0, 2 + 5 * C), np.float64)
    starts = np.flatnonzero(np.r_[True, lab_sorted[1:] != lab_sorted[:-1]])
    present = lab_sorted[starts]
    area = np.diff(np.r_[starts, lab_sorted.size])
    yy, xx = np.divmod(order, W)
    n = present.size
    ints = np.zeros((n, 6), np.int64)
    flts = np.zeros((n, 2 + 5 * C), np.float64)
    ints[:, 0] = present
    ints[:, 1] = area
    ints[:, 2] = np.minimum.reduceat(yy, starts)
    ints[:, 3] = np.minimum.reduceat(xx, starts)
    ints[:, 4] = np.maximum.reduceat(yy, starts) + 1
    ints[:, 5] = np.maximum.reduceat(xx, starts) + 1
    flts[:, 0] = np.add.reduceat(yy, starts) / area
    flts[:, 1] = np.add.reduceat(xx, starts) / area
    for c in range(C):
        v = vals[c].ravel()[order]
        o = 2 + 5 * c
        s = np.add.reduceat(v, starts)
        mean = s / area
        dev = v - np.repeat(mean, area)
        flts[:, o + 0] = s
        flts[:, o + 1] = mean
        flts[:, o + 2] = np.sqrt(np.add.reduceat(dev * dev, starts) / area)
        flts[:, o + 3] = np.minimum.reduceat(v, starts)
        flts[:, o + 4] = np.maximum.reduceat(v, starts)
    return ints, flts
