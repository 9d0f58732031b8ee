"""16-bit TIFF decode / encode for the drop-in scripts (Pillow; the reference uses imageio,
tifffile and Pillow for the same files: MaxProjection.py:39,48, Illumination_QC_mult.py:145,
Image_re-binning.py:17-21)."""
import io

import numpy as np
from PIL import Image


def decode(data):
    """bytes -> 2-D array (uint16 for 16-bit TIFFs, uint8 for 8-bit images)."""
    with Image.open(io.BytesIO(data)) as im:
        if im.mode in ("I;16", "I;16L", "I;16B"):
            return np.array(im, dtype=np.uint16)
        if im.mode == "L":
            return np.array(im, dtype=np.uint8)
        if im.mode in ("I", "F"):
            return np.array(im)
        raise ValueError("unsupported image mode %s" % im.mode)


def read(path):
    with open(path, "rb") as f:
        return decode(f.read())


def encode(arr, compression=None):
    """2-D uint16 / uint8 array -> TIFF bytes (compression None or 'tiff_lzw')."""
    arr = np.ascontiguousarray(arr)
    im = Image.fromarray(arr)
    buf = io.BytesIO()
    if compression:
        im.save(buf, format="tiff", compression=compression)
    else:
        im.save(buf, format="tiff")
    return buf.getvalue()


def write(path, arr, compression=None):
    with open(path, "wb") as f:
        f.write(encode(arr, compression))
