"""K7 parity: the device TIFF strip codec against Pillow/libtiff (files byte for byte, pixels
bit for bit), the oracle restatement and the reference's own output files."""
import io
import os
import struct

import numpy as np
import pytest
from PIL import Image

from oracle import tiff_lzw as T
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu


def pil_file(a, **kw):
    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="tiff", **kw)
    return buf.getvalue()


def stack(shape, seed, kind="noise"):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        a = np.clip(rng.normal(300, 30, shape), 0, 65535).astype(np.uint16)
    elif kind == "full":
        a = rng.integers(0, 65536, shape).astype(np.uint16)
    elif kind == "zeros":
        a = np.zeros(shape, np.uint16)
    else:
        a = (np.arange(int(np.prod(shape))) % 1000).reshape(shape).astype(np.uint16)
    return a


@pytest.mark.parametrize("shape,kind", [((3, 120, 200), "noise"), ((2, 300, 500), "zeros"), ((1, 90, 333), "full"),
                                        ((2, 300, 400), "ramp"), ((1, 1, 1), "noise"), ((2, 2000, 3), "noise"),
                                        ((1, 2, 40000), "ramp"), ((4, 67, 129), "full")])
def test_encode_equals_pillow_bytes(shape, kind):
    require_gpu()
    from image_processing_suite_b200.scripts import tiffio
    a = stack(shape, 3, kind)
    files = tiffio.encode_lzw_from_device(dev(a))
    assert len(files) == shape[0]
    for p in range(shape[0]):
        assert T.same_file(files[p], pil_file(a[p], compression="tiff_lzw"))
        assert files[p] == T.encode_tiff_lzw(a[p])


def test_encode_full_size_planes_equal_pillow():
    """Five 1080^2 planes (one re-binned field): 180 strips in one launch."""
    require_gpu()
    from image_processing_suite_b200.scripts import tiffio
    a = stack((5, 1080, 1080), 11)
    a[1] = stack((1080, 1080), 12, "zeros")
    a[2, 100:200, 100:900] = 65535
    files = tiffio.encode_lzw_from_device(dev(a))
    for p in range(5):
        assert T.same_file(files[p], pil_file(a[p], compression="tiff_lzw"))      # up to libtiff's unwritten pad byte


def test_encode_custom_strip_height_round_trips_through_pillow():
    require_gpu()
    from image_processing_suite_b200.scripts import tiffio
    a = stack((2, 100, 77), 4, "full")
    for rps in (1, 7, 100, 1000):
        for p, f in enumerate(tiffio.encode_lzw_from_device(dev(a), rows_per_strip=rps)):
            np.testing.assert_array_equal(np.asarray(Image.open(io.BytesIO(f)), dtype=np.uint16), a[p])


def big_endian_file(a, rps):
    """Uncompressed big-endian TIFF written by hand (Pillow only writes little-endian)."""
    h, w = a.shape
    body = bytearray(b"MM\x00*\x00\x00\x00\x00")
    offs, cnts = [], []
    for r in range(0, h, rps):
        offs.append(len(body))
        body += a[r:r + rps].astype(">u2").tobytes()
        cnts.append(len(body) - offs[-1])
    n = len(offs)
    ifd = len(body)
    tags = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, n, None),
            (278, 3, 1, rps), (279, 4, n, None)]
    after = ifd + 2 + 12 * len(tags) + 4
    out = struct.pack(">H", len(tags))
    for t, ty, c, v in tags:
        if v is None:
            arr = offs if t == 273 else cnts
            v = arr[0] if n == 1 else (after if t == 273 else after + 4 * n)
            out += struct.pack(">HHII", t, ty, c, v)
        elif ty == 3:
            out += struct.pack(">HHIHH", t, ty, c, v, 0)
        else:
            out += struct.pack(">HHII", t, ty, c, v)
    out += struct.pack(">I", 0)
    if n > 1:
        out += struct.pack(">%dI" % n, *offs) + struct.pack(">%dI" % n, *cnts)
    body += out
    body[4:8] = struct.pack(">I", ifd)
    return bytes(body)


def test_decode_every_supported_layout():
    require_gpu()
    from image_processing_suite_b200.scripts import tiffio
    a = stack((6, 150, 211), 5, "noise")
    a[3] = stack((150, 211), 6, "full")
    files = [pil_file(a[0], compression="tiff_lzw"), pil_file(a[1]), pil_file(a[2], compression="tiff_lzw", tiffinfo={317: 2}),
             pil_file(a[3], compression="tiff_lzw"), big_endian_file(a[4], 40), big_endian_file(a[5], 150)]
    np.testing.assert_array_equal(np.asarray(Image.open(io.BytesIO(files[4])), dtype=np.uint16), a[4])   # the hand-made file is valid
    got = host(tiffio.decode_to_device(files))
    np.testing.assert_array_equal(got, a)
    for f, ref in zip(files[:4], a[:4]):
        np.testing.assert_array_equal(T.decode_tiff(f), ref)


def test_decode_full_size_plane_single_strip_and_many():
    """2160^2: Pillow's 15-row strips and one 9.3 MB strip (writers that do not split)."""
    require_gpu()
    from image_processing_suite_b200 import ops
    from image_processing_suite_b200.scripts import tiffio
    a = stack((2160, 2160), 13)
    a[500:600, :] = 0
    f = pil_file(a, compression="tiff_lzw")
    np.testing.assert_array_equal(host(tiffio.decode_to_device([f]))[0], a)
    one = tiffio.encode_lzw_from_device(dev(a[None]), rows_per_strip=2160)[0]
    assert len(tiffio.parse(one)["offsets"]) == 1
    np.testing.assert_array_equal(host(tiffio.decode_to_device([one]))[0], a)
    np.testing.assert_array_equal(np.asarray(Image.open(io.BytesIO(one)), dtype=np.uint16), a)
    assert ops.tiff_rows_per_strip(2160, 2160) == 15 == T.rows_per_strip(2160, 2160)


def test_decode_reports_damaged_strips():
    require_gpu()
    import torch
    from image_processing_suite_b200 import ops
    from image_processing_suite_b200.scripts import tiffio
    a = stack((64, 64), 7)
    f = pil_file(a, compression="tiff_lzw")
    info = tiffio.parse(f)
    o, c = info["offsets"][0], info["counts"][0]
    src = torch.from_numpy(np.frombuffer(f, np.uint8).copy()).cuda()
    dst = torch.full((3 * 8192 + 5,), 0xAB, dtype=torch.uint8, device="cuda")
    st = ops.tiff_lzw_decode(src, [o, o, o], [c, c // 2, c], dst, [0, 8192, 16384 + 3], [8192, 8192, 100])
    assert st.tolist() == [0, 1, 0]
    d = host(dst)
    assert d[:8192].tobytes() == a.tobytes()
    assert d[16387:16487].tobytes() == a.tobytes()[:100] and (d[16487:] == 0xAB).all() and (d[16384:16387] == 0xAB).all()
    assert (d[8192 + 8000:16384] == 0).all()                      # the undecodable remainder is zero-filled
    bad = bytearray(f)
    bad[o + c // 2] ^= 0xFF
    bad[o + c // 2 + 1] ^= 0xFF
    try:
        got = host(tiffio.decode_to_device([bytes(bad)]))
        assert not np.array_equal(got[0], a)                       # a flipped byte either fails or changes pixels
    except ValueError:
        pass
    with pytest.raises(ValueError):
        ops.tiff_lzw_decode(src, [o], [len(f)], dst, [0], [8192])
    with pytest.raises(ValueError):
        tiffio.decode_to_device([f, pil_file(a[:32], compression="tiff_lzw")])


def test_fix_u16_matches_numpy():
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(2)
    for shape in ((3, 17, 1000), (5, 1), (2, 2160), (4, 257)):
        a = rng.integers(0, 65536, shape).astype(np.uint16)
        np.testing.assert_array_equal(host(ops.tiff_fix_u16(dev(a), 2, False)), np.cumsum(a, axis=-1, dtype=np.uint16))
        np.testing.assert_array_equal(host(ops.tiff_fix_u16(dev(a), 1, True)), a.byteswap())
        np.testing.assert_array_equal(host(ops.tiff_fix_u16(dev(a), 2, True)), np.cumsum(a.byteswap(), axis=-1, dtype=np.uint16))
        np.testing.assert_array_equal(host(ops.tiff_fix_u16(dev(a), 1, False)), a)


def test_rebinning_script_reproduces_reference_files(golden_dir):
    """Image_re-binning.process_image_in_memory, bytes in -> bytes out, against the files the
    reference function wrote for the same inputs (tests/golden/tiff_lzw.npz)."""
    require_gpu()
    from image_processing_suite_b200.scripts import Image_rebinning
    g = np.load(os.path.join(golden_dir, "tiff_lzw.npz"))
    for name in ("noise", "lzw_in", "flat", "tall", "ident"):
        ow, oh = (int(v) for v in g[f"{name}_size"])
        got = Image_rebinning.process_image_in_memory(g[f"{name}_in"].tobytes(), target_size=(ow, oh))
        assert T.same_file(got, g[f"{name}_file"].tobytes()), name


def test_rebinning_batch_and_host_decoded_inputs():
    require_gpu()
    from image_processing_suite_b200.scripts import Image_rebinning
    from oracle import lanczos as o_lz
    a = stack((3, 120, 160), 21, "noise")
    ins = [pil_file(a[0]), pil_file(a[1], compression="tiff_lzw"), pil_file(a[2], compression="tiff_adobe_deflate")]
    outs = Image_rebinning.process_images_in_memory(ins, (80, 60))          # deflate input: host decode, device encode
    for p in range(3):
        assert T.same_file(outs[p], pil_file(o_lz.pil_resize(a[p], (60, 80)), compression="tiff_lzw"))
    buf = io.BytesIO()
    Image.fromarray(a[0]).save(buf, format="png")
    assert Image_rebinning.process_image_in_memory(buf.getvalue(), (80, 60)) == outs[0]
    with pytest.raises(Exception):
        Image_rebinning.process_image_in_memory(b"not an image")


@pytest.mark.parametrize("shape", [(24, 97, 131), (12, 331, 1000)])
def test_codec_fuzz_against_pillow(shape):
    """Planes of different statistics in one launch: files equal Pillow's, pixels come back."""
    require_gpu()
    from image_processing_suite_b200.scripts import tiffio
    rng = np.random.default_rng(31)
    P, H, W = shape
    a = np.empty(shape, np.uint16)
    for p in range(P):
        kind = p % 6
        if kind == 0:
            a[p] = rng.integers(0, rng.integers(1, 4), (H, W))
        elif kind == 1:
            a[p] = np.clip(rng.normal(rng.integers(100, 3000), rng.integers(1, 200), (H, W)), 0, 65535)
        elif kind == 2:
            a[p] = np.repeat(rng.integers(0, 65536, (H, (W + 7) // 8)), 8, axis=1)[:, :W]
        elif kind == 3:
            a[p] = np.tile(rng.integers(0, 4096, rng.integers(2, 300)), H * W)[:H * W].reshape(H, W)
        elif kind == 4:
            a[p] = rng.integers(0, 65536, (H, W))
        else:
            a[p] = (np.add.outer(np.arange(H), np.arange(W)) // 3) % 65536
    files = tiffio.encode_lzw_from_device(dev(a))
    refs = [pil_file(a[p], compression="tiff_lzw") for p in range(P)]
    for p in range(P):
        assert T.same_file(files[p], refs[p]), p
    np.testing.assert_array_equal(host(tiffio.decode_to_device(refs)), a)
    pred = [pil_file(a[p], compression="tiff_lzw", tiffinfo={317: 2}) for p in range(P)]
    np.testing.assert_array_equal(host(tiffio.decode_to_device(pred)), a)


def test_entry_points_reject_bad_arguments():
    """Error behaviour of the C ABI: status codes, nothing written."""
    require_gpu()
    import ctypes as C
    import torch
    from image_processing_suite_b200 import capi
    H, W, rps = 64, 80, 7
    planes = torch.zeros((2, H, W), dtype=torch.uint16, device="cuda")
    cap = int(capi.call("ips_tiff_file_bound", H, W, rps))
    files = torch.empty((2, cap), dtype=torch.uint8, device="cuda")
    nbytes = torch.zeros((2,), dtype=torch.int64, device="cuda")
    ws = torch.empty((int(capi.call("ips_tiff_encode_workspace_bytes", 2, H, W, rps)),), dtype=torch.uint8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def code(*args):
        with pytest.raises(capi.IpsError) as e:
            capi.call("ips_tiff_lzw_encode_u16", *args)
        return e.value.code

    assert code(p(planes), 2, H, W, 0, p(files), cap, p(nbytes), p(ws), ws.numel(), st) == -1        # rows_per_strip
    assert code(p(planes), 2, H, W, rps, p(files), cap - 16, p(nbytes), p(ws), ws.numel(), st) == -6  # file_cap below the bound
    assert code(p(planes), 2, H, W, rps, p(files), cap, p(nbytes), p(ws), 16, st) == -6               # workspace
    assert code(p(planes), 2, H, W, rps, C.c_void_p(files.data_ptr() + 1), cap, p(nbytes), p(ws), ws.numel(), st) == -3
    assert code(C.c_void_p(0), 2, H, W, rps, p(files), cap, p(nbytes), p(ws), ws.numel(), st) == -7
    assert capi.call("ips_tiff_lzw_encode_u16", p(planes), 0, H, W, rps, p(files), cap, p(nbytes), p(ws), ws.numel(), st) == 0
    with pytest.raises(capi.IpsError) as e:
        capi.call("ips_tiff_fix_u16", p(planes), 2 * H, W, 3, 0, st)
    assert e.value.code == -7
    with pytest.raises(capi.IpsError) as e:
        capi.call("ips_tiff_lzw_decode", p(files), C.c_void_p(0), C.c_void_p(0), p(files), C.c_void_p(0), C.c_void_p(0), 1, C.c_void_p(0), st)
    assert e.value.code == -7
    assert int(capi.call("ips_tiff_rows_per_strip", 2160, 2160)) == 15 and int(capi.call("ips_tiff_rows_per_strip", 2, 40000)) == 1
    assert int(capi.call("ips_tiff_lzw_bound", 64800)) >= 64800 * 3 // 2
