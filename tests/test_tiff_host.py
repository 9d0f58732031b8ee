"""CPU checks of the TIFF strip codec: the oracle restatement against Pillow/libtiff and the
golden files the reference's process_image_in_memory wrote, the host build of the very code the CUDA
kernels run (csrc/tiff_lzw_core.cuh, 32 threads standing in for the lanes) against both, and
the product's TIFF parser."""
import ctypes
import io
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from oracle import lanczos as o_lz
from oracle import tiff_lzw as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pil_lzw(a, **kw):
    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="tiff", compression="tiff_lzw", **kw)
    return buf.getvalue()


def images():
    rng = np.random.default_rng(5)
    yield "noise", np.clip(rng.normal(300, 30, (120, 200)), 0, 65535).astype(np.uint16)
    yield "zeros", np.zeros((300, 500), np.uint16)              # ratio check resets the table early
    yield "const", np.full((64, 64), 65535, np.uint16)
    yield "full", rng.integers(0, 65536, (90, 333)).astype(np.uint16)   # table fills: reset at code 4094
    yield "ramp", (np.arange(400 * 300) % 1000).reshape(300, 400).astype(np.uint16)
    yield "one", np.array([[7]], np.uint16)
    yield "tall", rng.integers(0, 50, (2000, 3)).astype(np.uint16)
    yield "wide", rng.integers(0, 5, (2, 40000)).astype(np.uint16)      # stride > 64 KiB: one row per strip


@pytest.mark.parametrize("name,img", list(images()), ids=[n for n, _ in images()])
def test_oracle_files_equal_pillow(name, img):
    ref = pil_lzw(img)
    assert T.same_file(T.encode_tiff_lzw(img), ref)
    np.testing.assert_array_equal(T.decode_tiff(ref), img)


def test_same_file_ignores_only_the_pad_byte():
    img = np.random.default_rng(3).integers(0, 65536, (40, 50)).astype(np.uint16)
    for k in range(40):                                   # find an image whose strips end on an odd offset
        img[0, 0] = k
        f = bytearray(T.encode_tiff_lzw(img))
        info = T.parse_tiff(bytes(f))
        end = info["offsets"][-1] + info["counts"][-1]
        if end & 1:
            break
    assert end & 1
    g = bytearray(f)
    g[end] = 0xD7
    assert T.same_file(f, g) and T.same_file(g, f)
    g[end - 1] ^= 1
    assert not T.same_file(f, g)
    g = bytearray(f)
    g[end + 3] ^= 1
    assert not T.same_file(f, g)
    assert not T.same_file(f, f[:-1])


def test_oracle_reads_predictor_and_uncompressed():
    rng = np.random.default_rng(6)
    img = rng.integers(0, 65536, (50, 70)).astype(np.uint16)
    np.testing.assert_array_equal(T.decode_tiff(pil_lzw(img, tiffinfo={317: 2})), img)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="tiff")
    np.testing.assert_array_equal(T.decode_tiff(buf.getvalue()), img)


def test_oracle_reproduces_reference_files(golden_dir):
    """The reference's own output bytes (Image_re-binning.process_image_in_memory, run by
    oracle/make_golden.py) from the oracle's Lanczos + TIFF writer."""
    g = np.load(os.path.join(golden_dir, "tiff_lzw.npz"))
    for name in ("noise", "lzw_in", "flat", "tall", "ident"):
        src = T.decode_tiff(g[f"{name}_in"].tobytes())
        ow, oh = (int(v) for v in g[f"{name}_size"])
        assert T.same_file(T.encode_tiff_lzw(o_lz.pil_resize(src, (oh, ow))), g[f"{name}_file"].tobytes()), name


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lzw") / "lzw_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", os.path.join(ROOT, "tests", "native", "lzw_host_harness.cpp"), "-o", so],
                   check=True)
    h = ctypes.CDLL(so)
    h.harness_bound.restype = ctypes.c_uint64
    h.harness_bound.argtypes = [ctypes.c_uint64]
    h.harness_encode.restype = ctypes.c_uint32
    h.harness_encode.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
    h.harness_decode.restype = ctypes.c_int
    h.harness_decode.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
    return h


def _enc(h, raw, off=0):
    cap = h.harness_bound(len(raw))
    out = np.zeros(cap, np.uint8)
    src = np.zeros(len(raw) + 8, np.uint8)
    src[off:off + len(raw)] = np.frombuffer(raw, np.uint8)
    n = h.harness_encode(src.ctypes.data + off, len(raw), out.ctypes.data, cap)
    assert n != 0xFFFFFFFF
    return out[:n].tobytes()


def _dec(h, comp, n, off=0, ooff=0):
    out = np.full(n + 64, 0xAB, np.uint8)
    src = np.zeros(len(comp) + 8, np.uint8)
    src[off:off + len(comp)] = np.frombuffer(comp, np.uint8)
    st = h.harness_decode(src.ctypes.data + off, len(comp), out.ctypes.data + ooff, n)
    assert (out[:ooff] == 0xAB).all() and (out[ooff + n:] == 0xAB).all()      # nothing outside the strip
    return st, out[ooff:ooff + n].tobytes()


def harness_images():
    """Smaller than images(): 32 threads meeting at a barrier per warp collective are slow."""
    rng = np.random.default_rng(15)
    yield "noise", np.clip(rng.normal(300, 30, (24, 200)), 0, 65535).astype(np.uint16)
    yield "zeros", np.zeros((70, 500), np.uint16)               # two strips; ratio check resets the table early
    yield "const", np.full((64, 64), 65535, np.uint16)
    yield "full", rng.integers(0, 65536, (40, 333)).astype(np.uint16)    # table fills: reset at code 4094
    yield "ramp", (np.arange(100 * 300) % 1000).reshape(100, 300).astype(np.uint16)
    yield "one", np.array([[7]], np.uint16)
    yield "tall", rng.integers(0, 50, (900, 3)).astype(np.uint16)
    yield "wide", rng.integers(0, 5, (2, 33000)).astype(np.uint16)      # stride > 64 KiB: one row per strip


@pytest.mark.parametrize("name,img", list(harness_images()), ids=[n for n, _ in harness_images()])
def test_kernel_state_machines_equal_pillow(harness, name, img):
    """Every strip Pillow writes: the kernels' encoder reproduces its bytes and the kernels'
    decoder its pixels, at every input / output alignment."""
    info = T.parse_tiff(pil_lzw(img))
    ref = pil_lzw(img)
    rps = info["rps"]
    for s, (o, c) in enumerate(zip(info["offsets"], info["counts"])):
        raw = img[s * rps:(s + 1) * rps].tobytes()
        for off in ((0, 1, 2, 3) if s == 0 and len(raw) < 12000 else ((s + 1) % 4, (s + 3) % 4)):   # every alignment on small first strips
            assert _enc(harness, raw, off) == ref[o:o + c]
            st, px = _dec(harness, ref[o:o + c], len(raw), off, (5 * off + s) % 16)
            assert st == 0 and px == raw


def test_kernel_decoder_reports_damage(harness):
    rng = np.random.default_rng(8)
    raw = np.clip(rng.normal(300, 30, 10000), 0, 65535).astype(np.uint16).tobytes()
    comp = _enc(harness, raw)
    st, px = _dec(harness, comp[:len(comp) // 2], len(raw))
    assert st == 1 and px[-100:] == bytes(100)                       # truncated: remainder zero-filled
    st, px = _dec(harness, comp, 1000)                               # fewer bytes wanted than coded: fine
    assert st == 0 and px == raw[:1000]
    assert _dec(harness, b"\x00\x01" + comp, 100)[0] == 3            # pre-6.0 bit order
    assert _dec(harness, bytes(rng.integers(0, 256, 500, dtype=np.uint8)), 4000)[0] in (1, 2)
    assert _dec(harness, b"", 16)[0] == 1


def test_kernel_encoder_overflow_is_reported(harness):
    raw = np.random.default_rng(9).integers(0, 256, 4096, dtype=np.uint8).tobytes()
    out = np.zeros(1024, np.uint8)
    src = np.frombuffer(raw, np.uint8).copy()
    assert harness.harness_encode(src.ctypes.data, len(raw), out.ctypes.data, 1024) == 0xFFFFFFFF


def test_product_parser_matches_oracle_and_rejects_other_layouts():
    from image_processing_suite_b200.scripts import tiffio
    img = np.arange(100 * 40, dtype=np.uint16).reshape(100, 40)
    for kw in ({}, {"compression": "tiff_lzw"}, {"compression": "tiff_lzw", "tiffinfo": {317: 2}}):
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="tiff", **kw)
        a, b = tiffio.parse(buf.getvalue()), T.parse_tiff(buf.getvalue())
        assert (a["width"], a["height"], a["compression"], a["predictor"], a["rows_per_strip"]) == \
               (b["width"], b["height"], b["compression"], b["predictor"], b["rps"])
        assert a["offsets"] == b["offsets"] and a["counts"] == b["counts"]
        # the same tables as int64 arrays (what decode_staged takes, built on the reader threads)
        assert a["offsets_np"].dtype == np.int64 and a["offsets_np"].tolist() == a["offsets"]
        assert a["counts_np"].dtype == np.int64 and a["counts_np"].tolist() == a["counts"]
    for kw in ({"compression": "tiff_adobe_deflate"},):
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="tiff", **kw)
        with pytest.raises(tiffio.Unsupported):
            tiffio.parse(buf.getvalue())
    buf = io.BytesIO()
    Image.fromarray(img.astype(np.uint8)).save(buf, format="tiff")
    with pytest.raises(tiffio.Unsupported):
        tiffio.parse(buf.getvalue())
    buf = io.BytesIO()
    Image.fromarray(img.astype(np.uint8)).save(buf, format="png")
    with pytest.raises(tiffio.Unsupported):
        tiffio.parse(buf.getvalue())
    with pytest.raises(tiffio.Unsupported):
        tiffio.parse(b"II*\x00\xff\xff\xff\x7f")


def _fuzz_strip(rng, kind, n):
    if kind == 0:      # few symbols: long strings, references far beyond the decoder's output window
        return rng.integers(0, rng.integers(1, 4), n, dtype=np.uint8)
    if kind == 1:      # 16-bit samples with a quiet high byte, like a camera image
        return np.clip(rng.normal(rng.integers(100, 3000), rng.integers(1, 200), n // 2), 0, 65535).astype("<u2").view(np.uint8)
    if kind == 2:      # runs of random length
        out = np.repeat(rng.integers(0, 256, n // 4 + 1, dtype=np.uint8), rng.integers(1, 40, n // 4 + 1))
        return out[:n]
    if kind == 3:      # periodic with noise
        base = np.tile(rng.integers(0, 256, rng.integers(2, 700), dtype=np.uint8), n)[:n].copy()
        flip = rng.random(n) < 0.01
        base[flip] = rng.integers(0, 256, int(flip.sum()), dtype=np.uint8)
        return base
    return rng.integers(0, 256, n, dtype=np.uint8)   # incompressible


@pytest.mark.parametrize("seed", range(10))
def test_kernel_state_machines_fuzz(harness, seed):
    """Random strips of five entropy classes and odd lengths: the kernels' encoder equals the
    oracle's libtiff restatement, the kernels' decoder returns the bytes, truncated requests too."""
    rng = np.random.default_rng(1000 + seed)
    kind = seed % 5
    n = int(rng.integers(1, 40000))
    raw = np.ascontiguousarray(_fuzz_strip(rng, kind, n)).tobytes()
    off = int(rng.integers(0, 4))
    comp = _enc(harness, raw, off)
    assert comp == T.lzw_encode_strip(raw)
    st, px = _dec(harness, comp, len(raw), int(rng.integers(0, 4)), int(rng.integers(0, 16)))
    assert st == 0 and px == raw
    cut = int(rng.integers(1, len(raw) + 1))
    st, px = _dec(harness, comp, cut, 0, int(rng.integers(0, 16)))
    assert st == 0 and px == raw[:cut]


def test_oracle_fuzz_against_pillow():
    """Seeded sweep of small images of every entropy class and odd shapes: the oracle's files equal
    Pillow's (up to the pad byte) and decode back."""
    rng = np.random.default_rng(77)
    for trial in range(60):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 120))
        kind = trial % 5
        raw = np.ascontiguousarray(_fuzz_strip(rng, kind, 2 * h * w + 8))[:2 * h * w]
        img = raw.view("<u2").reshape(h, w).astype(np.uint16)
        ref = pil_lzw(img)
        assert T.same_file(T.encode_tiff_lzw(img), ref), (trial, h, w)
        np.testing.assert_array_equal(T.decode_tiff(ref), img)


def test_plain_writer_is_a_valid_baseline_tiff():
    """tiffio.encode without compression (the output of the max-projection script): one strip,
    readable by Pillow, by the oracle and by the product parser; odd sizes pad to an even IFD offset."""
    from image_processing_suite_b200.scripts import tiffio
    rng = np.random.default_rng(4)
    for shape in ((1, 1), (3, 5), (7, 9), (64, 33), (200, 300)):
        img = rng.integers(0, 65536, shape).astype(np.uint16)
        data = tiffio.encode(img)
        np.testing.assert_array_equal(np.asarray(Image.open(io.BytesIO(data)), dtype=np.uint16), img)
        np.testing.assert_array_equal(T.decode_tiff(data), img)
        info = tiffio.parse(data)
        assert (info["width"], info["height"], info["compression"], len(info["offsets"])) == (shape[1], shape[0], 1, 1)
        assert info["counts"][0] == img.nbytes and int.from_bytes(data[4:8], "little") % 2 == 0
    with pytest.raises(tiffio.Unsupported):
        tiffio.parse(data[:len(data) // 2])                 # IFD gone
    cut = bytearray(tiffio.encode(img, "tiff_lzw"))
    info = tiffio.parse(bytes(cut))
    ifd = int.from_bytes(cut[4:8], "little")
    with pytest.raises(tiffio.Unsupported):                  # a strip table that points outside the file
        broken = bytes(cut[:ifd]) + bytes(cut[ifd:]).replace(int(info["counts"][0]).to_bytes(4, "little"), (1 << 30).to_bytes(4, "little"), 1)
        tiffio.parse(broken)
