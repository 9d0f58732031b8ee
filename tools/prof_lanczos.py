"""K5 timing: Pillow-exact LANCZOS of one 5-plane 2160^2 field to 1080^2, 540^2 and 720^2 (register
kernels for integer decimation) and to 1000^2 (generic staged kernels).  One JSON line each."""
import json
import sys

import torch

sys.path.insert(0, ".")
from image_processing_suite_b200 import ops

PEAK = 6451.8
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device="cuda").manual_seed(0)
fields = [torch.randint(200, 4000, (5, 2160, 2160), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16) for _ in range(12)]
for res in (1080, 540, 720, 1000):
    for _ in range(4):
        ops.lanczos_resize_u16(fields[0], (res, res))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(48):
        ops.lanczos_resize_u16(fields[i % 12], (res, res))       # 12 distinct fields = 560 MB: larger than L2
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 48
    nbytes = 5 * (2160 * 2160 + res * res) * 2
    print(json.dumps({"kernel": "lanczos 2160 -> %d, 5 planes" % res, "ms": ms, "algorithmic_bytes": nbytes,
                      "gbs": nbytes / ms / 1e6, "frac_of_measured_hbm": nbytes / ms / 1e6 / PEAK}), flush=True)
