"""Oracle: centroid-centred masked cell crops and their 8-bit scaling.

Test infrastructure only -- see oracle/__init__.py.

Restates the crop loop of Cellpose_GPU_s3fs.py:149-170 (inside ``consumer_worker``, which
cannot be imported without cellpose / transformers) and ``scale_to_8bit`` (:34-43, which can:
oracle/make_golden.py runs the reference's own function for tests/golden/crops.npz).
``regionprops(...).centroid`` is the mean of the object's pixel coordinates (float64);
scikit-image is not installed here, the mean is taken directly.
"""
import numpy as np


def scale_to_8bit(image):
    """Cellpose_GPU_s3fs.py:34-43: min-max scale to [0, 255], truncate to uint8; a constant
    image becomes zeros."""
    min_val, max_val = np.min(image), np.max(image)
    if max_val == min_val:
        return np.zeros(image.shape, dtype=np.uint8)
    scaled = 255.0 * (image.astype(np.float32) - min_val) / (max_val - min_val)
    return scaled.astype(np.uint8)


def cell_crops(image_hwc, masks, box=200):
    """Cellpose_GPU_s3fs.py:149-182 for one field.

    image_hwc [H][W][C] float32 (illumination-corrected channels stacked last, :72-73),
    masks [H][W] integer labels.  For every object in ascending label order: integer-truncated
    centroid, dropped when the box leaves the image (:159-161), crop masked by
    ``mask == label`` (:163-166), every channel scaled on its own (:176-178).
    Returns (crops uint8 [n][C][box][box], coords int [n][2] (y, x), labels int [n]).
    """
    image_hwc = np.asarray(image_hwc)
    masks = np.asarray(masks)
    h, w, c = image_hwc.shape
    half = box // 2
    labels = np.unique(masks)
    labels = labels[labels > 0]
    crops, coords, kept = [], [], []
    for lab in labels:
        ys, xs = np.nonzero(masks == lab)
        yc, xc = int(ys.mean()), int(xs.mean())
        if yc - half < 0 or yc + half > h or xc - half < 0 or xc + half > w:
            continue
        y1, y2, x1, x2 = yc - half, yc + half, xc - half, xc + half
        binary = (masks[y1:y2, x1:x2] == lab)[:, :, None]
        crop = image_hwc[y1:y2, x1:x2, :] * binary
        crops.append(np.stack([scale_to_8bit(crop[:, :, k]) for k in range(c)]))
        coords.append((yc, xc))
        kept.append(int(lab))
    if not crops:
        return (np.zeros((0, c, 2 * half, 2 * half), np.uint8), np.zeros((0, 2), np.int64), np.zeros((0,), np.int64))
    return np.stack(crops), np.asarray(coords, np.int64), np.asarray(kept, np.int64)
