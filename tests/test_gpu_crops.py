"""Cell crops (CUDA) vs the restated crop loop of Cellpose_GPU_s3fs.py and the reference's
own scale_to_8bit goldens.  Bit-exact (uint8 pixels, integer centroids, kept labels)."""
import os

import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import crops as o_crops
from oracle import preprocess as o_pre
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu


def test_scale_to_8bit_golden_through_the_crop_kernel(golden_dir):
    """An all-labelled (box+1)^2 canvas has its centroid at box/2, so its crop is canvas[:box, :box]
    unmasked -- exactly scale_to_8bit of the golden image placed there: checked against the
    reference function's outputs."""
    require_gpu()
    from image_processing_suite_b200 import ops
    g = np.load(os.path.join(golden_dir, "crops.npz"))
    assert int(g["box_size"]) == 200
    for name in ("a", "const"):
        img = g[f"{name}_in"]
        box = img.shape[0]
        canvas = np.zeros((box + 1, box + 1), np.float32)
        canvas[:box, :box] = img
        lab = np.ones((1, box + 1, box + 1), np.int32)
        ints = np.array([[[1, (box + 1) ** 2, 0, 0, box + 1, box + 1]]], np.int32)
        r = ops.cell_crops(dev(canvas[None, None]), dev(lab), dev(ints), dev(np.array([1], np.int32)), box=box)
        assert host(r["kept"])[0, 0].tolist() == [1, box // 2, box // 2]
        assert int(host(r["n_kept"])[0]) == 1
        np.testing.assert_array_equal(host(r["crops"])[0, 0, 0], g[f"{name}_out"])
        np.testing.assert_array_equal(o_crops.scale_to_8bit(img), g[f"{name}_out"])


@pytest.mark.parametrize("box", [40, 26])
def test_cell_crops_match_restated_loop(box):
    require_gpu()
    from image_processing_suite_b200 import ops
    F, C, H, W, cells = 2, 3, 160, 208, 24
    labs = np.stack([synth.make_labels(H, W, cells, seed=30 + f, amin=5, amax=11) for f in range(F)])
    mps = np.stack([o_pre.max_projection_field(synth.field_numpy(labs[f], c=C, z=2, seed=f)) for f in range(F)])
    ill = synth.make_illum(C, H, W, seed=2)
    k1 = ops.preprocess_fused(dev(mps[:, :, None]), dev(ill), bin=1, want_corrected=True, want_binned=False)
    k3 = ops.object_stats(dev(labs), k1["maxproj"], None, 1.0, n_max=cells)
    r = ops.cell_crops(k1["corrected"], dev(labs), k3["ints"], k3["n_objects"], box=box)
    corr = host(k1["corrected"])                                   # the kernel's own float32 quotients
    for f in range(F):
        e_crops, e_coords, e_labels = o_crops.cell_crops(np.moveaxis(corr[f], 0, -1), labs[f], box)
        n = int(host(r["n_kept"])[f])
        assert n == e_crops.shape[0] and 0 < n < cells              # some cells sit on the edge
        kept = host(r["kept"])[f, :n]
        np.testing.assert_array_equal(kept[:, 0], e_labels)
        np.testing.assert_array_equal(kept[:, 1:], e_coords)
        np.testing.assert_array_equal(host(r["crops"])[f, :n], e_crops)


def test_cell_crops_capacity_and_errors():
    torch = require_gpu()
    from image_processing_suite_b200 import capi, ops
    H = W = 64
    lab = np.zeros((1, H, W), np.int32)
    lab[0, 20:24, 20:24] = 1
    lab[0, 40:44, 30:34] = 2
    lab[0, 0:3, 0:3] = 3                                            # on the edge: dropped
    img = np.random.default_rng(0).random((1, 1, H, W)).astype(np.float32)
    k3 = ops.object_stats(dev(lab), dev((img[:, :] * 1000).astype(np.uint16)), None, 1.0, n_max=4)
    r = ops.cell_crops(dev(img), dev(lab), k3["ints"], k3["n_objects"], box=16, max_crops=1)
    assert int(host(r["n_kept"])[0]) == 2                           # two pass the edge test, one slot written
    assert host(r["kept"])[0, 0].tolist() == [1, 21, 21]
    with pytest.raises(capi.IpsError):
        ops.cell_crops(dev(img), dev(lab), k3["ints"], k3["n_objects"], box=15)       # odd box
    with pytest.raises(capi.IpsError):
        ops.cell_crops(dev(img), dev(lab), k3["ints"], k3["n_objects"], box=128)      # larger than the image
