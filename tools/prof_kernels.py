"""Small driver for ncu captures: a few launches of K1 / K3 / fused on a 16-field batch."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from image_processing_suite_b200 import ops, synth

which = sys.argv[1] if len(sys.argv) > 1 else "k1k3"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
masks = bench.ring_masks(F, bench.H_, bench.W_, bench.NCELLS, seed=0)
labels = torch.empty((F, bench.H_, bench.W_), dtype=torch.int32, device=dev)
raw = torch.empty((F, bench.C_, bench.Z_, bench.H_, bench.W_), dtype=torch.uint16, device=dev)
for i, m in enumerate(masks):
    labels[i].copy_(torch.from_numpy(m))
    raw[i].copy_(synth.field_torch(labels[i], c=bench.C_, z=bench.Z_, seed=i))
illum = torch.from_numpy(synth.make_illum(bench.C_, bench.H_, bench.W_, seed=0)).to(dev)
k1 = ops.preprocess_fused(raw, illum, bin=2)
k3 = ops.object_stats(labels, k1["maxproj"], illum, 1 / 65535.0, n_max=bench.NCELLS)
fz = None
labels16 = labels.to(torch.uint16)
illum_rcp = ops.illum_reciprocal(illum)
for it in range(4):
    if "k1" in which:
        ops.preprocess_fused(raw, illum, bin=2, out=k1)
    if "k3" in which:
        ops.object_stats(labels, k1["maxproj"], illum, 1 / 65535.0, n_max=bench.NCELLS, out=k3)
    if "fused" in which:
        # as bench.py runs it: uint16 label masks, the function as its reciprocal
        fz = ops.field_fused(raw, illum, labels16, bin=2, intensity_scale=1 / 65535.0, n_max=bench.NCELLS, out=fz,
                             illum_rcp=illum_rcp)
torch.cuda.synchronize()
print("ok", which, F)
