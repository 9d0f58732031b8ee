"""MaxProjection.py on files: one field = 5 channels x 3 planes of 2160^2 as LZW TIFF bytes ->
5 max-projected planes.  GPU path of the drop-in script (strips decoded on the device, fused
z-max kernel) next to the reference's arithmetic on one host core (Pillow/libtiff decode +
np.maximum.reduce, MaxProjection.py:39-45; its TIFF writing is not timed).  One JSON line."""
import io
import json
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, ".")
from image_processing_suite_b200 import synth
from image_processing_suite_b200.scripts import MaxProjection, tiffio

C, Z = 5, 3
labs = synth.make_labels(2160, 2160, 2000, seed=5)
raw = synth.field_numpy(labs, c=C, z=Z, seed=5)                         # [C][Z][2160][2160]
files = {}
for c in range(C):
    for z in range(Z):
        b = io.BytesIO()
        Image.fromarray(raw[c, z]).save(b, format="tiff", compression="tiff_lzw")
        files[f"c{c}z{z}.tiff"] = b.getvalue()
ratio = sum(len(v) for v in files.values()) / raw.nbytes


class Mem:                                                                # the two calls the script makes on its client
    def __init__(self):
        self.out = {}

    def get_object(self, Bucket, Key):
        return {"Body": io.BytesIO(files[Key])}

    def upload_fileobj(self, f, Bucket, Key):
        self.out[Key] = f.read()


GROUPS = [[f"c{c}z{z}.tiff" for z in range(Z)] for c in range(C)]


def gpu_field():
    mem = Mem()
    assert MaxProjection.max_project_chunk(GROUPS, "b", mem) == C            # fetch, decode, project, encode, upload
    return mem.out


gpu_field()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    got = gpu_field()
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / 3

t0 = time.perf_counter()
ref = np.maximum.reduce([np.asarray(Image.open(io.BytesIO(files[f"c0z{z}.tiff"]))) for z in range(Z)])
t_cpu = (time.perf_counter() - t0) * C
assert np.array_equal(tiffio.decode(got[MaxProjection.modify_imagepath('c0z0.tiff')]), ref)
print(json.dumps({"step": "MaxProjection.max_project_chunk of one field: 15 LZW TIFF planes of 2160^2 in, 5 TIFF planes out (host bytes both ways)",
                  "lzw_bytes_over_pixel_bytes": ratio, "gpu_ms_per_field": t_gpu * 1e3, "cpu_reference_ms_per_field_1core": t_cpu * 1e3,
                  "fields_per_s_gpu": 1 / t_gpu, "fields_per_s_cpu_1core": 1 / t_cpu, "results_match": True}))
