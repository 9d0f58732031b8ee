"""K7 on content of different compressibility: 80 planes of 1080^2 per class, device-timed
encode and decode (strips resident), LZW bytes over pixel bytes.  One JSON line per class."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from image_processing_suite_b200 import ops
from image_processing_suite_b200.scripts import tiffio


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:1080, 0:1080]
smooth = 300 + 2000 * np.exp(-((yy - 500) ** 2 + (xx - 600) ** 2) / 2e5)
classes = {
    "camera noise sigma 30 (incompressible)": lambda: np.clip(smooth + rng.normal(0, 30, smooth.shape), 0, 65535),
    "camera noise sigma 3": lambda: np.clip(smooth + rng.normal(0, 3, smooth.shape), 0, 65535),
    "noise-free smooth signal": lambda: smooth,
    "mostly background (90 % constant)": lambda: np.where(rng.random(smooth.shape) < 0.1, smooth + rng.normal(0, 30, smooth.shape), 300),
}
for name, make in classes.items():
    planes = torch.from_numpy(np.stack([make().astype(np.uint16) for _ in range(4)])).cuda().repeat(20, 1, 1).contiguous()
    px = planes.numel() * 2
    ms_enc = timed(lambda: ops.tiff_lzw_encode(planes))
    files, nbytes = ops.tiff_lzw_encode(planes)
    blobs = [bytes(files[p, :int(n)].cpu().numpy()) for p, n in enumerate(nbytes)]
    infos = [tiffio.parse(b) for b in blobs]
    base, tot = [], 0
    for b in blobs:
        base.append(tot)
        tot += (len(b) + 15) // 16 * 16
    host = np.zeros(tot, np.uint8)
    for b0, b in zip(base, blobs):
        host[b0:b0 + len(b)] = np.frombuffer(b, np.uint8)
    src = torch.from_numpy(host).cuda()
    dst = torch.empty(px, dtype=torch.uint8, device="cuda")
    so = np.concatenate([b0 + np.asarray(i["offsets"]) for b0, i in zip(base, infos)])
    sb = np.concatenate([np.asarray(i["counts"]) for i in infos])
    rps = infos[0]["rows_per_strip"]
    ns = len(infos[0]["offsets"])
    rows = np.minimum(rps, 1080 - rps * np.arange(ns))
    do = np.concatenate([p * 1080 * 1080 * 2 + rps * 2160 * np.arange(ns) for p in range(len(infos))])
    db = np.tile(rows * 2160, len(infos))
    ms_dec = timed(lambda: ops.tiff_lzw_decode(src, so, sb, dst, do, db))
    assert torch.equal(dst.view(torch.uint16).reshape(planes.shape), planes)
    print(json.dumps({"content": name, "lzw_over_pixels": float(nbytes.sum()) / px, "encode_ms": ms_enc, "encode_pixel_gbs": px / ms_enc / 1e6,
                      "decode_ms": ms_dec, "decode_pixel_gbs": px / ms_dec / 1e6}), flush=True)
