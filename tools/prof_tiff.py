"""One encode and one decode launch of the TIFF-LZW codec on 5 planes of 1080^2 (for ncu)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from image_processing_suite_b200.scripts import tiffio

rng = np.random.default_rng(0)
a = np.clip(rng.normal(300, 30, (5, 1080, 1080)), 0, 65535).astype(np.uint16)
d = torch.from_numpy(a).cuda()
for _ in range(3):
    files = tiffio.encode_lzw_from_device(d)
    back = tiffio.decode_to_device(files)
torch.cuda.synchronize()
assert torch.equal(back, d)
print("ok", sum(len(f) for f in files))
