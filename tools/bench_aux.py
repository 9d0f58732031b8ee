"""Config-3 style measurements of the secondary kernels: illumination accumulate / finalize /
median, Pillow-exact Lanczos re-binning, K1 bin sweep.  One JSON line per kernel."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from image_processing_suite_b200 import ops

PEAK = 6451.8
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=20, warm=3):
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


def line(name, ms, nbytes, **kw):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac_of_measured_hbm": gbs / PEAK, **kw}), flush=True)


C, H, W = 5, 2160, 2160
g = torch.Generator(device="cuda").manual_seed(0)
F = 64
fields = torch.randint(200, 4000, (F, C, H, W), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)   # 3 GB

# K2 accumulate: 16 fields per call, cycling over 64 resident fields (> L2)
est = ops.IllumEstimator(C, H, W)
state = {"i": 0}
def acc():
    i = state["i"] % (F // 16)
    state["i"] += 1
    est.n = 0
    est.add(fields[i * 16:(i + 1) * 16])
ms = timed(acc)
line("illum_accumulate_kernel (16 fields/call)", ms, 16 * C * H * W * 2 + 2 * C * H * W * 4, fields_per_s=16 / (ms * 1e-3))
est.n = 64
for sigma in (20.0, 85.0):
    ms = timed(lambda: est.finalize(sigma), iters=3, warm=1)
    line("illum_finalize (mean + gaussian sigma=%g + select + rescale), per plate" % sigma, ms, 6 * C * H * W * 4, sigma=sigma)
ms = timed(lambda: ops.illum_median(fields[:32]), iters=3, warm=1)
line("illum_median_kernel (N=32)", ms, 16 * 32 * C * H * W * 2, note="16 bisection passes over the stack")

# K5 Lanczos: one field (5 planes) per call
for res in (1080, 540):
    st = {"i": 0}
    def lz():
        i = st["i"] % F
        st["i"] += 1
        ops.lanczos_resize_u16(fields[i], (res, res))
    ms = timed(lz)
    line("lanczos 2160->%d (h + v pass, 5 planes/call)" % res, ms, C * (H * W * 2 + 2 * H * res * 2 + res * res * 2),
         fields_per_s=1 / (ms * 1e-3))

# CPU reference for Lanczos (Pillow, 1 thread, 1 plane)
from PIL import Image
plane = fields[0, 0].cpu().numpy()
t0 = time.perf_counter()
for _ in range(3):
    Image.fromarray(plane).resize((1080, 1080), resample=Image.Resampling.LANCZOS)
print(json.dumps({"kernel": "Pillow LANCZOS 2160->1080, 1 plane, 1 thread", "ms": (time.perf_counter() - t0) / 3 * 1e3}), flush=True)

# K1 bin sweep (with function), 16 fields per call from a 32-field raw ring
raw = torch.randint(0, 65536, (32, C, 3, H, W), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
ill = 1.0 + torch.rand((C, H, W), device="cuda", generator=g) * 0.5
for b in (1, 2, 4):
    out = ops.preprocess_fused(raw[:16], ill, bin=b)
    st = {"i": 0}
    def k1():
        i = st["i"] % 2
        st["i"] += 1
        ops.preprocess_fused(raw[i * 16:(i + 1) * 16], ill, bin=b, out=out)
    ms = timed(k1)
    nbytes = 16 * (C * 3 * H * W * 2 + C * H * W * 2 + C * (H // b) * (W // b) * 4) + C * H * W * 4
    line("preprocess_vec_kernel bin=%d (16 fields/call)" % b, ms, nbytes, fields_per_s=16 / (ms * 1e-3))
