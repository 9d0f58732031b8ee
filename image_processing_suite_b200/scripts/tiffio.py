"""16-bit TIFF decode / encode for the drop-in scripts (the reference uses imageio, tifffile and
Pillow for the same files: MaxProjection.py:39,48, Illumination_QC_mult.py:145,
Image_re-binning.py:17-21).

Strips of 16-bit single-sample TIFFs (uncompressed or LZW, predictor 1 or 2, either byte order)
are decoded and LZW-encoded on the GPU (decode_to_device / encode_lzw_from_device / load_planes,
K7); decode() / encode() are the host side for everything else the reference accepts (PNG, JPEG,
tiled or deflate TIFFs) and for the plain uncompressed output of the max-projection script."""
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


def plain_u16_parts(h, w):
    """(head, tail) of the uncompressed little-endian baseline TIFF of an h x w uint16 image: the
    file is head + pixels + tail (header, the pixels as one strip, one IFD -- what
    imageio.imwrite(..., format='tiff') amounts to, MaxProjection.py:48)."""
    import struct
    nbytes = h * w * 2
    pad = nbytes & 1
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, 8),
            (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, nbytes)]
    ifd = struct.pack("<H", len(tags)) + b"".join(struct.pack("<HHII", *t) for t in tags) + struct.pack("<I", 0)
    return b"II*\x00" + struct.pack("<I", 8 + nbytes + pad), b"\x00" * pad + ifd


def _encode_plain_u16(arr):
    h, w = arr.shape
    data = np.ascontiguousarray(arr, dtype="<u2")
    head, tail = plain_u16_parts(h, w)
    return b"".join([head, memoryview(data).cast("B"), tail])


def encode(arr, compression=None):
    """2-D uint16 / uint8 array -> TIFF bytes (compression None or 'tiff_lzw')."""
    arr = np.ascontiguousarray(arr)
    if not compression and arr.dtype == np.uint16 and arr.ndim == 2 and arr.nbytes < 2 ** 32 - 1024:
        return _encode_plain_u16(arr)
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


# ---- device codec (K7) ---------------------------------------------------------------------
# The strips of 16-bit single-sample TIFFs are decoded and LZW-encoded on the GPU
# (ips_tiff_lzw_decode / ips_tiff_lzw_encode_u16): compressed bytes are what crosses PCIe.
# Anything outside that layout (tiles, other compressions, PNG/JPEG inputs the reference also
# accepts, Image_re-binning.py:40) raises Unsupported and the caller uses decode() above.
class Unsupported(ValueError):
    pass


_TYPE_FMT = {1: "B", 3: "H", 4: "I", 16: "Q"}


def parse(data):
    """First IFD of a classic TIFF -> dict; raises Unsupported for anything the device codec
    does not read."""
    import struct
    if len(data) < 8 or data[:2] not in (b"II", b"MM"):
        raise Unsupported("not a TIFF")
    e = "<" if data[:2] == b"II" else ">"
    magic, ifd = struct.unpack_from(e + "HI", data, 2)
    if magic != 42:
        raise Unsupported("BigTIFF / unknown magic %d" % magic)
    if ifd + 2 > len(data):
        raise Unsupported("IFD outside the file")
    (n,) = struct.unpack_from(e + "H", data, ifd)
    if ifd + 2 + 12 * n > len(data):
        raise Unsupported("IFD outside the file")
    tags = {}
    for k in range(n):
        tag, typ, cnt = struct.unpack_from(e + "HHI", data, ifd + 2 + 12 * k)
        fmt = _TYPE_FMT.get(typ)
        if fmt is None:
            continue
        size = struct.calcsize(fmt) * cnt
        at = ifd + 10 + 12 * k
        if size > 4:
            (at,) = struct.unpack_from(e + "I", data, at)
        if at + size > len(data):
            raise Unsupported("tag %d outside the file" % tag)
        tags[tag] = struct.unpack_from(e + "%d%s" % (cnt, fmt), data, at)
    if 256 not in tags or 257 not in tags or 273 not in tags or 279 not in tags:
        raise Unsupported("not a stripped TIFF")
    if 322 in tags or 324 in tags:
        raise Unsupported("tiled TIFF")
    w, h = int(tags[256][0]), int(tags[257][0])
    info = dict(width=w, height=h, big_endian=(e == ">"), bits=tags.get(258, (1,))[0],
                samples=tags.get(277, (1,))[0], compression=tags.get(259, (1,))[0],
                predictor=tags.get(317, (1,))[0], offsets=list(tags[273]), counts=list(tags[279]),
                rows_per_strip=min(int(tags.get(278, (h,))[0]), h), sample_format=tags.get(339, (1,))[0])
    if info["bits"] != 16 or info["samples"] != 1 or info["sample_format"] != 1:
        raise Unsupported("device codec reads 16-bit unsigned single-sample images")
    if info["compression"] not in (1, 5) or info["predictor"] not in (1, 2):
        raise Unsupported("compression %d / predictor %d" % (info["compression"], info["predictor"]))
    if w <= 0 or h <= 0 or info["rows_per_strip"] <= 0:
        raise Unsupported("empty image")
    n_strips = -(-h // info["rows_per_strip"])
    if len(info["offsets"]) != n_strips or len(info["counts"]) != n_strips:
        raise Unsupported("strip tables do not match the image height")
    # the strip tables once more as int64 arrays: built here, on the reader thread, so that the plate loop's
    # launching thread does not convert thousands of Python integers per batch (decode_staged)
    info["offsets_np"] = np.asarray(info["offsets"], np.int64)
    info["counts_np"] = np.asarray(info["counts"], np.int64)
    if np.any(info["offsets_np"] + info["counts_np"] > len(data)):
        raise Unsupported("strip outside the file")
    return info


def decode_staged(src, infos, bases, out=None, defer_check=False):
    """The device half of the codec: ``src`` is a uint8 CUDA tensor that holds whole TIFF files,
    file p at byte ``bases[p]`` (16-byte aligned) with the IFD facts ``infos[p]`` (``parse``).
    Returns the uint16 CUDA tensor [P][H][W] (``out`` if given).  LZW strips are decoded by
    ips_tiff_lzw_decode, uncompressed strips are device-to-device copies; compressed bytes are all
    that ever crossed PCIe.  ValueError when shapes differ or a strip is corrupt.
    ``defer_check``: do not wait for the decoder here; returns (out, check) and the caller runs
    ``check()`` once the stream has been synchronised (a pipelined loop must not block per batch)."""
    import torch
    from .. import ops
    if not infos:
        raise ValueError("no files")
    h, w = infos[0]["height"], infos[0]["width"]
    if any((i["height"], i["width"]) != (h, w) for i in infos):
        raise ValueError("Image shape mismatch")
    dev = src.device
    plane = h * w * 2
    if out is None:
        out = torch.empty((len(infos), h, w), dtype=torch.uint16, device=dev)
    dst = out.view(torch.uint8).reshape(-1)
    so, sb, do, db = [], [], [], []
    for p, i in enumerate(infos):
        rps = i["rows_per_strip"]
        offs, cnts = i.get("offsets_np"), i.get("counts_np")
        if offs is None or cnts is None:
            offs, cnts = np.asarray(i["offsets"], np.int64), np.asarray(i["counts"], np.int64)
        rows = np.minimum(rps, h - rps * np.arange(len(offs), dtype=np.int64))
        d0, dn = p * plane + rps * w * 2 * np.arange(len(offs), dtype=np.int64), rows * w * 2
        if i["compression"] == 5:
            so.append(int(bases[p]) + offs)
            sb.append(cnts)
            do.append(d0)
            db.append(dn)
            continue
        if np.any(cnts < dn):
            raise ValueError("an uncompressed strip of file %d is short" % p)
        if np.all(offs[1:] == offs[:-1] + dn[:-1]):                 # the usual case: one run of pixels
            o = int(bases[p]) + int(offs[0])
            dst[p * plane:(p + 1) * plane].copy_(src[o:o + plane], non_blocking=True)
        else:
            for o, a, n in zip(offs, d0, dn):
                dst[int(a):int(a + n)].copy_(src[int(bases[p]) + int(o):int(bases[p]) + int(o + n)], non_blocking=True)
    status = None
    if so:
        so, sb, do, db = (np.concatenate(x) for x in (so, sb, do, db))
        status = ops.tiff_lzw_decode(src, so, sb, dst, do, db, to_host=not defer_check)

    def check():
        if status is None:
            return
        st = status.cpu() if status.is_cuda else status
        if int(st.max()) != 0:
            bad = int(torch.nonzero(st)[0])
            if int(st[bad]) == 3:
                raise Unsupported("pre-6.0 (LSB-first) LZW strip")          # libtiff still reads those: host decode
            raise ValueError("LZW strip %d is damaged (status %d)" % (bad, int(st[bad])))
    if not defer_check:
        check()
    for p, i in enumerate(infos):
        if i["big_endian"] or i["predictor"] == 2:
            ops.tiff_fix_u16(out[p], predictor=i["predictor"], byteswap=i["big_endian"])
    return (out, check) if defer_check else out


def decode_to_device(files, device=None):
    """list of TIFF byte strings of equal shape -> uint16 CUDA tensor [P][H][W].
    Raises Unsupported (before touching the GPU) when any file is outside the device codec's
    layout, ValueError when shapes differ or a strip is corrupt."""
    import warnings
    import torch
    infos = [parse(f) for f in files]
    if not infos:
        raise ValueError("no files")
    if any((i["height"], i["width"]) != (infos[0]["height"], infos[0]["width"]) for i in infos):
        raise ValueError("Image shape mismatch")
    dev = torch.device(device if device is not None else "cuda")
    bases, total = [], 0
    for f in files:
        bases.append(total)
        total += (len(f) + 15) // 16 * 16
    src = torch.empty((max(total, 16),), dtype=torch.uint8, device=dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)              # torch.frombuffer on read-only bytes
        for f, b in zip(files, bases):    # bytes -> device, straight from the bytes object (the driver stages it once)
            src[b:b + len(f)].copy_(torch.frombuffer(f, dtype=torch.uint8))
    return decode_staged(src, infos, bases)


def encode_lzw_from_device(planes, rows_per_strip=None):
    """uint16 CUDA tensor [P][H][W] -> list of P LZW TIFF byte strings, the bytes Pillow's
    save(format='TIFF', compression='tiff_lzw') writes for the same pixels."""
    from .. import ops
    files, nbytes = ops.tiff_lzw_encode(planes, rows_per_strip)
    if not len(nbytes):
        return []
    host = files[:, :int(nbytes.max())].cpu().numpy()
    return [host[p, :int(n)].tobytes() for p, n in enumerate(nbytes)]


def load_planes(files, device=None):
    """list of encoded 16-bit images -> uint16 CUDA tensor [P][H][W]: the device codec when every
    file is a TIFF it reads, otherwise decode() on the host and one copy.  ValueError when the
    planes differ in shape or are not 16-bit."""
    import logging
    import torch
    try:
        return decode_to_device(files, device)
    except Unsupported as why:
        # not a silent switch: the reference decodes these with PIL too (PNG / JPEG / tiled or deflate TIFFs)
        logging.getLogger(__name__).info("host image decoder used (%s); the arithmetic still runs on the GPU", why)
    images = [decode(f) for f in files]
    if not all(img.shape == images[0].shape for img in images):
        raise ValueError("Image shape mismatch")
    if any(img.dtype != np.uint16 or img.ndim != 2 for img in images):
        raise ValueError("expected single-plane 16-bit images, got %s %s" % (images[0].dtype, images[0].shape))
    return torch.from_numpy(np.ascontiguousarray(np.stack(images))).to(device if device is not None else "cuda")
