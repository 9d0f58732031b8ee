"""K7 measurements: device TIFF-LZW encode / decode of re-binned fields (5 x 1080^2 uint16 per
field), Pillow/libtiff on one host core beside it, and the whole Image_re-binning step
(TIFF bytes -> LANCZOS -> LZW TIFF bytes) through the script's function.  One JSON line each."""
import io
import json
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, ".")
from image_processing_suite_b200 import ops, synth
from image_processing_suite_b200.scripts import Image_rebinning, tiffio


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def out(**kw):
    print(json.dumps(kw), flush=True)


F, C = 16, 5
labs = synth.make_labels(2160, 2160, 2000, seed=1)
raw = np.stack([synth.field_numpy(labs, c=C, z=1, seed=f)[:, 0] for f in range(2)])          # [2][C][2160][2160]
big = torch.from_numpy(raw).cuda().repeat(F // 2, 1, 1, 1).reshape(F * C, 2160, 2160).contiguous()
small = ops.lanczos_resize_u16(big, (1080, 1080))                                              # [80][1080][1080]
P = small.shape[0]
px_bytes = P * 1080 * 1080 * 2

# encode kernels only (no D2H)
ms = timed(lambda: ops.tiff_lzw_encode(small))
files, nbytes = ops.tiff_lzw_encode(small)
out(kernel="tiff_lzw_encode (80 planes 1080^2, 2880 strips)", ms=ms, fields_per_s=F / (ms * 1e-3), pixel_gbs=px_bytes / ms / 1e6,
    compressed_fraction=float(nbytes.sum()) / px_bytes)
ms1 = timed(lambda: ops.tiff_lzw_encode(small[:1]))
out(kernel="tiff_lzw_encode (1 plane 1080^2, 36 strips: latency)", ms=ms1)

# decode kernels only: strips already on the device
blobs = tiffio.encode_lzw_from_device(small)
infos = [tiffio.parse(b) for b in blobs]
base, tot = [], 0
for b in blobs:
    base.append(tot)
    tot += (len(b) + 15) // 16 * 16
hostbuf = np.zeros(tot, np.uint8)
for b0, b in zip(base, blobs):
    hostbuf[b0:b0 + len(b)] = np.frombuffer(b, np.uint8)
src = torch.from_numpy(hostbuf).cuda()
dst = torch.empty(px_bytes, dtype=torch.uint8, device="cuda")
so, sb, do, db = [], [], [], []
for p, (b0, i) in enumerate(zip(base, infos)):
    for s, (o, c) in enumerate(zip(i["offsets"], i["counts"])):
        rows = min(i["rows_per_strip"], 1080 - s * i["rows_per_strip"])
        so.append(b0 + o); sb.append(c); do.append(p * 1080 * 1080 * 2 + s * i["rows_per_strip"] * 2160); db.append(rows * 2160)
import ctypes as Ct
from image_processing_suite_b200 import capi
d64 = torch.tensor([so, do], dtype=torch.int64).cuda()
d32 = torch.tensor([sb, db], dtype=torch.int64).to(torch.uint32).cuda()
status = torch.empty(len(so), dtype=torch.int32, device="cuda")
def dec():
    capi.call("ips_tiff_lzw_decode", Ct.c_void_p(src.data_ptr()), Ct.c_void_p(d64[0].data_ptr()), Ct.c_void_p(d32[0].data_ptr()),
              Ct.c_void_p(dst.data_ptr()), Ct.c_void_p(d64[1].data_ptr()), Ct.c_void_p(d32[1].data_ptr()), len(so),
              Ct.c_void_p(status.data_ptr()), Ct.c_void_p(torch.cuda.current_stream().cuda_stream))
ms = timed(dec)
assert int(status.max()) == 0 and torch.equal(dst.view(torch.uint16).reshape(P, 1080, 1080), small)
out(kernel="tiff_lzw_decode (80 planes 1080^2, 2880 strips)", ms=ms, fields_per_s=F / (ms * 1e-3), pixel_gbs=px_bytes / ms / 1e6)

# Pillow / libtiff on one host core, same planes
hs = small[:10].cpu().numpy()
t0 = time.perf_counter()
ref = []
for a in hs:
    b = io.BytesIO()
    Image.fromarray(a).save(b, format="tiff", compression="tiff_lzw")
    ref.append(b.getvalue())
t_enc = (time.perf_counter() - t0) / len(hs)
t0 = time.perf_counter()
for b in ref:
    np.asarray(Image.open(io.BytesIO(b)))
t_dec = (time.perf_counter() - t0) / len(hs)
assert all(len(a) == len(b) and sum(x != y for x, y in zip(a, b)) <= 1 for a, b in zip(ref, blobs[:10]))   # pad byte
out(cpu="Pillow/libtiff, 1 core", encode_ms_per_plane=t_enc * 1e3, decode_ms_per_plane=t_dec * 1e3,
    encode_fields_per_s=1 / (t_enc * C), decode_fields_per_s=1 / (t_dec * C), files_identical=True)

# the whole re-binning step through the script function: 16 LZW TIFFs of 2160^2 in -> 16 LZW TIFFs of 1080^2 out
ins = tiffio.encode_lzw_from_device(big[:16])
t0 = time.perf_counter()
outs = Image_rebinning.process_images_in_memory(ins, (1080, 1080))
torch.cuda.synchronize()
t1 = time.perf_counter()
outs = Image_rebinning.process_images_in_memory(ins, (1080, 1080))
t_gpu = time.perf_counter() - t1
t0 = time.perf_counter()
for b in ins[:4]:
    im = Image.open(io.BytesIO(b)).resize((1080, 1080), resample=Image.Resampling.LANCZOS)
    o = io.BytesIO()
    im.save(o, format="TIFF", compression="tiff_lzw")
    assert len(o.getvalue()) == len(outs[ins.index(b)])
t_cpu = (time.perf_counter() - t0) / 4
out(step="Image_re-binning.process_image_in_memory, bytes -> bytes (2160^2 LZW -> 1080^2 LZW)", gpu_ms_per_image=t_gpu / 16 * 1e3,
    cpu_ms_per_image=t_cpu * 1e3, images_per_s_gpu=16 / t_gpu, images_per_s_cpu_1core=1 / t_cpu, files_identical=True)
