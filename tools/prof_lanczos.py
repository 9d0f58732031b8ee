import sys, torch
sys.path.insert(0, ".")
from image_processing_suite_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randint(200, 4000, (5, 2160, 2160), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
for _ in range(4):
    ops.lanczos_resize_u16(x, (1080, 1080))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.lanczos_resize_u16(x, (1080, 1080))
e1.record(); torch.cuda.synchronize()
print("ms per call", e0.elapsed_time(e1) / 20)
