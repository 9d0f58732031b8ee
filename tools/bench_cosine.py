"""Throughput of the tensor-core cosine path on a config-5-shaped group (scaled rows)."""
import json
import sys
import torch
sys.path.insert(0, ".")
from image_processing_suite_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, d), device="cuda", generator=g)
for _ in range(2):
    s, npairs = ops.cosine_triu(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
s, npairs = ops.cosine_triu(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
xh = x.double() / x.double().norm(dim=1, keepdim=True)
ref = 0.5 * (float((xh.sum(0) ** 2).sum()) - float((xh * xh).sum()))
flops = n * (n - 1) / 2 * 2 * d
print(json.dumps({"n": n, "d": d, "ms": ms, "algorithmic_tflops": flops / ms / 1e9,
                  "issued_bf16_tflops": 3 * (n / 128) * (n / 128 + 1) / 2 * 128 * 128 * 2 * ((d + 63) // 64 * 64) / ms / 1e9,
                  "mean_cos": float(s[0]) / int(npairs[0]), "mean_cos_ref": ref / int(npairs[0]),
                  "abs_err_mean": abs(float(s[0]) - ref) / int(npairs[0])}))
