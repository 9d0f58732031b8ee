"""K3 parity: per-object statistics (CUDA, through the C-ABI) vs the scipy.ndimage oracle.

Bit-exact: object count, labels, area, half-open bbox.  Float columns (centroid, sum,
mean, std, min, max): RTOL = 1e-5 relative (north_star), std additionally with an
absolute floor of 1e-5 * mean (std of a near-constant object is a cancellation result).
"""
import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import object_stats as o_obj
from oracle import preprocess as o_pre
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _compare(res, f, lab, mp, ill, scale, C):
    n = int(host(res["n_objects"])[f])
    e_i, e_f = o_obj.object_stats(lab, mp, ill, scale)
    assert n == e_i.shape[0]
    np.testing.assert_array_equal(host(res["ints"])[f, :n], e_i)
    got = host(res["flts"])[f, :n].astype(np.float64)
    np.testing.assert_allclose(got[:, :2], e_f[:, :2], rtol=RTOL, atol=1e-6)
    for c in range(C):
        o = 2 + 5 * c
        for k in (0, 1, 3, 4):
            np.testing.assert_allclose(got[:, o + k], e_f[:, o + k], rtol=RTOL, atol=0)
        np.testing.assert_allclose(got[:, o + 2], e_f[:, o + 2], rtol=RTOL,
                                   atol=RTOL * np.abs(e_f[:, o + 1]).max())


@pytest.mark.parametrize("shape,cells", [((2, 3, 128, 160), 30), ((1, 5, 96, 256), 20), ((3, 1, 64, 64), 6)])
@pytest.mark.parametrize("with_illum", [False, True])
def test_object_stats_matches_oracle(shape, cells, with_illum):
    require_gpu()
    from image_processing_suite_b200 import ops
    F, C, H, W = shape
    labs = np.stack([synth.make_labels(H, W, cells, seed=40 + f, amin=5, amax=12) for f in range(F)])
    mps = np.stack([o_pre.max_projection_field(synth.field_numpy(labs[f], c=C, z=2, seed=f)) for f in range(F)])
    ill = synth.make_illum(C, H, W, seed=1) if with_illum else None
    scale = 1.0 / 65535.0 if with_illum else 1.0
    res = ops.object_stats(dev(labs), dev(mps), dev(ill) if with_illum else None, scale, n_max=cells + 3)
    for f in range(F):
        _compare(res, f, labs[f], mps[f], ill, scale, C)


def test_object_stats_sparse_and_absent_labels():
    """Absent labels are skipped, rows ascend by label, single-pixel and full-width objects."""
    require_gpu()
    from image_processing_suite_b200 import ops
    H, W = 40, 48
    lab = np.zeros((1, H, W), np.int32)
    lab[0, 2:5, 3:7] = 2
    lab[0, 6, 0] = 5
    lab[0, 10, :] = 9                        # one full row
    lab[0, 12:30, 47] = 7                    # last column
    lab[0, 39, 47] = 11                      # corner pixel
    lab[0, 20:22, 10:30:2] = 3               # non-contiguous object (alternating columns)
    rng = np.random.default_rng(0)
    mp = rng.integers(0, 65536, (1, 2, H, W), dtype=np.uint16)
    res = ops.object_stats(dev(lab), dev(mp), None, 1.0, n_max=16)
    _compare(res, 0, lab[0], mp[0], None, 1.0, 2)
    np.testing.assert_array_equal(host(res["ints"])[0, :6, 0], [2, 3, 5, 7, 9, 11])


def test_object_stats_empty_mask_and_overflow_flag():
    require_gpu()
    from image_processing_suite_b200 import ops
    lab = np.zeros((2, 32, 32), np.int32)
    lab[1, 4:8, 4:8] = 50                    # label above n_max
    mp = np.ones((2, 1, 32, 32), np.uint16)
    res = ops.object_stats(dev(lab), dev(mp), None, 1.0, n_max=10)
    assert host(res["n_objects"]).tolist() == [0, -1]


def test_object_stats_constant_object_has_zero_std():
    require_gpu()
    from image_processing_suite_b200 import ops
    lab = np.zeros((1, 64, 64), np.int32)
    lab[0, 8:40, 8:40] = 1
    mp = np.full((1, 1, 64, 64), 40000, np.uint16)
    res = ops.object_stats(dev(lab), dev(mp), None, 1.0, n_max=1)
    row = host(res["flts"])[0, 0]
    assert row[2 + 2] == 0.0 and row[2 + 1] == 40000.0 and row[2 + 0] == 40000.0 * 1024


def test_object_stats_ragged_width():
    require_gpu()
    from image_processing_suite_b200 import ops
    H, W = 50, 77
    lab = synth.make_labels(H, W, 8, seed=9, amin=4, amax=8)[None]
    rng = np.random.default_rng(1)
    mp = rng.integers(0, 65536, (1, 3, H, W), dtype=np.uint16)
    ill = (1.0 + rng.random((3, H, W))).astype(np.float32)
    res = ops.object_stats(dev(lab), dev(mp), dev(ill), 1.0, n_max=8)
    _compare(res, 0, lab[0], mp[0], ill, 1.0, 3)


def test_object_stats_full_size_properties():
    """Config-2 size: areas sum to the foreground pixel count, every bbox contains its
    centroid, sum == mean * area, min <= mean <= max; the faster sort-based oracle
    restatement checks the integer columns exactly on one field."""
    torch = require_gpu()
    from image_processing_suite_b200 import ops
    H = W = 2160
    lab = synth.make_labels(H, W, 2000, seed=123)
    labs = dev(np.stack([lab, np.roll(lab, 7, axis=1)]))
    g = torch.Generator(device="cuda").manual_seed(2)
    mp = torch.randint(0, 65536, (2, 5, H, W), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
    res = ops.object_stats(labs, mp, None, 1.0, n_max=2000)
    n = host(res["n_objects"])
    ints, flts = host(res["ints"]), host(res["flts"]).astype(np.float64)
    assert n[0] == lab.max() and n[1] == lab.max()
    for f in range(2):
        I, Fl = ints[f, :n[f]], flts[f, :n[f]]
        assert I[:, 1].sum() == np.count_nonzero(lab)
        assert (I[:, 2] <= Fl[:, 0]).all() and (Fl[:, 0] < I[:, 4]).all()
        assert (I[:, 3] <= Fl[:, 1]).all() and (Fl[:, 1] < I[:, 5]).all()
        for c in range(5):
            o = 2 + 5 * c
            np.testing.assert_allclose(Fl[:, o], Fl[:, o + 1] * I[:, 1], rtol=1e-6)
            assert (Fl[:, o + 3] <= Fl[:, o + 1] + 1e-3).all() and (Fl[:, o + 1] <= Fl[:, o + 4] + 1e-3).all()
    e_i, e_f = o_obj.object_stats_bincount(lab, host(mp[0]), None, 1.0)
    np.testing.assert_array_equal(ints[0, :n[0]], e_i)
    np.testing.assert_allclose(flts[0, :n[0]], e_f, rtol=RTOL, atol=1e-6)
