"""CPU oracle for the TIFF strip codec at the edge of Image_re-binning.py / MaxProjection.py.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.

The reference writes its outputs with ``img.save(buf, format='TIFF', compression='tiff_lzw')``
(Image_re-binning.py:19-21) and reads planes with imageio / tifffile / PIL
(MaxProjection.py:39, Illumination_QC_mult.py:145, Image_re-binning.py:17).  The codec itself
lives in libtiff (bundled in Pillow; 4.7.1 with the Pillow 12.2.0 of this image), which is not
under /root/reference; this file restates its published LZW variant (TIFF 6.0 section 13 with
libtiff's encoder policy: MSB-first codes, 9..12 bits, "early change", ClearCode at the start of
every strip, table reset when code 4093 has been assigned, ratio check every 10000 input bytes)
and Pillow's file layout (strips first, IFD after them, byte counts then offsets after the IFD).
Pinned: tests/test_tiff_host.py compares every byte of the files this module writes with the
ones Pillow writes and with the files the reference's process_image_in_memory returned
(tests/golden/tiff_lzw.npz), and the pixels this module decodes with the ones Pillow decodes.

One byte is outside the contract: when the strips end on an odd offset libtiff seeks past one
pad byte to word-align the IFD and never writes it; Pillow's in-memory sink leaves it as
whatever its realloc'ed buffer held (observed 0x00 in one process and 0xd7 in another for the
same image; MALLOC_PERTURB_ changes it).  This writer and the device writer store 0 there and
same_file() ignores that byte.
"""
import struct

import numpy as np

BITS_MIN, BITS_MAX = 9, 12
CODE_CLEAR, CODE_EOI, CODE_FIRST = 256, 257, 258
CODE_MAX = (1 << BITS_MAX) - 1
CHECK_GAP = 10000
STRIP_SIZE = 65536        # Pillow: rows per strip = min(STRIP_SIZE // stride, height), at least 1


def same_file(a, b):
    """True when two TIFF byte strings are identical up to the unwritten pad byte before the IFD."""
    a, b = bytes(a), bytes(b)
    if a == b:
        return True
    if len(a) != len(b) or len(b) < 8:
        return False
    try:
        info = parse_tiff(b)
    except (KeyError, ValueError, struct.error):
        return False
    end = max(o + c for o, c in zip(info["offsets"], info["counts"]))
    ifd = struct.unpack(info["byteorder"] + "I", b[4:8])[0]
    if end & 1 and ifd == end + 1:
        return a[:end] == b[:end] and a[end + 1:] == b[end + 1:]
    return False


def rows_per_strip(width, height, bytes_per_pixel=2):
    stride = width * bytes_per_pixel
    return max(1, min(STRIP_SIZE // stride if stride else 1, height))


def lzw_encode_strip(data):
    """bytes of one strip -> LZW bytes, as libtiff's LZWPreEncode/LZWEncode/LZWPostEncode."""
    out = bytearray()
    nextdata = 0
    nextbits = 0
    nbits = BITS_MIN
    maxcode = (1 << nbits) - 1
    free_ent = CODE_FIRST
    checkpoint = CHECK_GAP
    ratio = 0
    incount = 0
    outcount = 0
    table = {}

    def put(code):
        nonlocal nextdata, nextbits, outcount
        nextdata = ((nextdata << nbits) | code) & 0xFFFFFFFFFFFF
        nextbits += nbits
        while nextbits >= 8:
            out.append((nextdata >> (nextbits - 8)) & 0xFF)
            nextbits -= 8
        outcount += nbits

    data = bytes(data)
    n = len(data)
    if n == 0:
        put(CODE_EOI)
        if nextbits:
            out.append((nextdata << (8 - nextbits)) & 0xFF)
        return bytes(out)
    put(CODE_CLEAR)
    ent = data[0]
    incount = 1
    for i in range(1, n):
        c = data[i]
        incount += 1
        key = (c << BITS_MAX) + ent
        hit = table.get(key)
        if hit is not None:
            ent = hit
            continue
        put(ent)
        ent = c
        table[key] = free_ent
        free_ent += 1
        if free_ent == CODE_MAX - 1:
            table.clear()
            ratio = 0
            incount = 0
            outcount = 0
            free_ent = CODE_FIRST
            put(CODE_CLEAR)
            nbits = BITS_MIN
            maxcode = (1 << nbits) - 1
        elif free_ent > maxcode:
            nbits += 1
            maxcode = (1 << nbits) - 1
        elif incount >= checkpoint:
            checkpoint = incount + CHECK_GAP
            if incount > 0x007FFFFF:
                rat = outcount >> 8
                rat = 0x7FFFFFFF if rat == 0 else incount // rat
            else:
                rat = (incount << 8) // outcount
            if rat <= ratio:
                table.clear()
                ratio = 0
                incount = 0
                outcount = 0
                free_ent = CODE_FIRST
                put(CODE_CLEAR)
                nbits = BITS_MIN
                maxcode = (1 << nbits) - 1
            else:
                ratio = rat
    # LZWPostEncode
    put(ent)
    free_ent += 1
    if free_ent == CODE_MAX - 1:
        outcount = 0
        put(CODE_CLEAR)
        nbits = BITS_MIN
    elif free_ent > maxcode:
        nbits += 1
    put(CODE_EOI)
    if nextbits:
        out.append((nextdata << (8 - nextbits)) & 0xFF)
    return bytes(out)


def lzw_decode_strip(comp, expected):
    """LZW bytes of one strip -> ``expected`` decoded bytes (libtiff LZWDecode, MSB-first)."""
    comp = bytes(comp)
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    nbits = BITS_MIN
    bitpos = 0
    total = len(comp) * 8
    old = None
    while len(out) < expected:
        if bitpos + nbits > total:
            break
        v = 0
        for k in range(nbits):            # small strips only: this is a checker, not a codec
            p = bitpos + k
            v = (v << 1) | ((comp[p >> 3] >> (7 - (p & 7))) & 1)
        bitpos += nbits
        if v == CODE_EOI:
            break
        if v == CODE_CLEAR:
            table = table[:258]
            nbits = BITS_MIN
            old = None
            continue
        if old is None:
            s = table[v]
        elif v < len(table):
            s = table[v]
            table.append(old + s[:1])
        else:
            s = old + old[:1]
            table.append(s)
        out += s
        old = s
        if len(table) + 1 > (1 << nbits) - 1 and nbits < BITS_MAX:
            nbits += 1
    return bytes(out[:expected])


def encode_tiff_lzw(img):
    """2-D uint16 array -> the bytes Pillow's ``save(format='tiff', compression='tiff_lzw')`` writes."""
    img = np.ascontiguousarray(img, dtype="<u2")
    h, w = img.shape
    rps = rows_per_strip(w, h)
    strips = [lzw_encode_strip(img[r:r + rps].tobytes()) for r in range(0, h, rps)]
    return assemble_tiff(w, h, rps, strips)


def assemble_tiff(w, h, rps, strips, compression=5):
    n = len(strips)
    body = bytearray(b"II*\x00\x00\x00\x00\x00")
    offsets = []
    for s in strips:
        offsets.append(len(body))
        body += s
    if len(body) & 1:
        body += b"\x00"
    ifd_off = len(body)
    counts = [len(s) for s in strips]
    tags = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, 16), (259, 3, 1, compression), (262, 3, 1, 1),
            (273, 4, n, None), (278, 3, 1, rps), (279, 4, n, None), (284, 3, 1, 1)]
    after = ifd_off + 2 + 12 * len(tags) + 4
    counts_off = after
    offsets_off = after + (4 * n if n > 1 else 0)
    ifd = struct.pack("<H", len(tags))
    for t, ty, c, v in tags:
        if t == 273:
            v = offsets[0] if n == 1 else offsets_off
        elif t == 279:
            v = counts[0] if n == 1 else counts_off
        ifd += struct.pack("<HHII", t, ty, c, v)
    ifd += struct.pack("<I", 0)
    body += ifd
    if n > 1:
        body += struct.pack("<%dI" % n, *counts)
        body += struct.pack("<%dI" % n, *offsets)
    body[4:8] = struct.pack("<I", ifd_off)
    return bytes(body)


TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8}


def parse_tiff(data):
    """First IFD of a classic TIFF -> dict(width, height, bits, compression, predictor, rps,
    offsets, counts, byteorder, samples, photometric, planar).  Raises ValueError on anything else."""
    data = bytes(data)
    if len(data) < 8 or data[:2] not in (b"II", b"MM"):
        raise ValueError("not a TIFF")
    bo = "<" if data[:2] == b"II" else ">"
    if struct.unpack(bo + "H", data[2:4])[0] != 42:
        raise ValueError("not a classic TIFF")
    off = struct.unpack(bo + "I", data[4:8])[0]
    n = struct.unpack(bo + "H", data[off:off + 2])[0]
    tags = {}
    for i in range(n):
        e = data[off + 2 + 12 * i: off + 14 + 12 * i]
        t, ty, c = struct.unpack(bo + "HHI", e[:8])
        size = TYPE_SIZE.get(ty, 1) * c
        raw = e[8:12] if size <= 4 else data[struct.unpack(bo + "I", e[8:12])[0]:][:size]
        fmt = {3: "H", 4: "I", 1: "B", 16: "Q"}.get(ty)
        if fmt is None:
            continue
        tags[t] = list(struct.unpack(bo + "%d%s" % (c, fmt), raw[:size]))
    w, h = tags[256][0], tags[257][0]
    return dict(width=w, height=h, bits=tags.get(258, [1])[0], compression=tags.get(259, [1])[0],
                predictor=tags.get(317, [1])[0], rps=min(tags.get(278, [h])[0], h), offsets=tags[273],
                counts=tags[279], byteorder=bo, samples=tags.get(277, [1])[0],
                photometric=tags.get(262, [1])[0], planar=tags.get(284, [1])[0])


def decode_tiff(data):
    """Classic single-sample 16-bit TIFF (uncompressed or LZW, predictor 1 or 2) -> (H,W) uint16."""
    info = parse_tiff(data)
    if info["bits"] != 16 or info["samples"] != 1:
        raise ValueError("16-bit single-sample TIFFs only")
    w, h, rps = info["width"], info["height"], info["rps"]
    out = np.empty((h, w), np.uint16)
    for s, (o, c) in enumerate(zip(info["offsets"], info["counts"])):
        r0 = s * rps
        rows = min(rps, h - r0)
        nbytes = rows * w * 2
        if info["compression"] == 1:
            raw = data[o:o + nbytes]
        elif info["compression"] == 5:
            raw = lzw_decode_strip(data[o:o + c], nbytes)
        else:
            raise ValueError("compression %d" % info["compression"])
        px = np.frombuffer(raw, dtype=info["byteorder"] + "u2").reshape(rows, w).astype(np.uint16)
        if info["predictor"] == 2:
            px = np.cumsum(px, axis=1, dtype=np.uint16)
        out[r0:r0 + rows] = px
    return out
