import sys, torch
sys.path.insert(0, ".")
from image_processing_suite_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randint(200, 4000, (5, 2160, 2160), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
for _ in range(3):
    ops.lanczos_resize_u16(x, (1080, 1080))
torch.cuda.synchronize()
print("ok")
