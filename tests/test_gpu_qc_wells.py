"""K6 ring sums and well aggregation (CUDA) vs the oracle."""
import numpy as np
import pytest
import scipy.fft
import scipy.ndimage

from oracle import normalize as o_norm
from oracle import qc as o_qc
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(64, 64), (96, 80), (45, 130), (216, 216)])
def test_ring_sums_match_oracle(hw):
    require_gpu()
    import torch
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(2)
    imgs = rng.random((2,) + hw) * 1000.0
    labels = o_qc.ring_labels(*hw)
    spec = torch.from_numpy(np.stack([scipy.fft.fft2(im - im.mean()) for im in imgs])).cuda()
    mag, pw = ops.ring_sums(spec, labels.size)
    for f in range(2):
        a = np.abs(scipy.fft.fft2(imgs[f] - imgs[f].mean()))
        rings = o_qc.ring_index(*hw)
        np.testing.assert_allclose(host(mag)[f], scipy.ndimage.sum(a, rings, labels), rtol=1e-10)
        np.testing.assert_allclose(host(pw)[f], scipy.ndimage.sum(a ** 2, rings, labels), rtol=1e-10)


def test_well_mean_matches_pandas_groupby():
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(5)
    n_wells, D = 24, 33
    wells = np.sort(rng.integers(0, n_wells, 5000)).astype(np.int32)
    wells[wells == 7] = 8                                   # well 7 stays empty
    rows = rng.normal(100.0, 30.0, (5000, D)).astype(np.float32)
    mean, count = ops.well_mean(dev(rows), dev(wells), n_wells)
    ids, ref = o_norm.well_mean(rows, wells)
    m, c = host(mean), host(count)
    np.testing.assert_allclose(m[ids], ref, rtol=1e-12)
    assert np.isnan(m[7]).all() and c[7] == 0
    np.testing.assert_array_equal(c, np.bincount(wells, minlength=n_wells))
    # unsorted ids give the same means
    perm = rng.permutation(5000)
    mean2, _ = ops.well_mean(dev(rows[perm]), dev(wells[perm]), n_wells)
    np.testing.assert_allclose(host(mean2)[ids], ref, rtol=1e-12)


def test_mad_robustize_and_double_sigmoid_match_oracle():
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(8)
    W, D = 384, 57
    prof = rng.normal(50.0, 9.0, (W, D))
    prof[5] = np.nan                                         # an empty well (no rows -> NaN means)
    ctrl = np.zeros(W, bool)
    ctrl[rng.choice(W, 33, replace=False)] = True
    ctrl[5] = True                                           # NaN control is ignored (nanmedian)
    prof[:, 3] = 7.0                                         # constant column: MAD 0 -> division by eps
    z = host(ops.mad_robustize(dev(prof), dev(ctrl.astype(np.uint8))))
    ref = o_norm.mad_robustize(prof, ctrl)
    ok = np.isfinite(ref)
    np.testing.assert_allclose(z[ok], ref[ok], rtol=1e-12, atol=1e-12)
    assert np.isnan(np.delete(z[5], 3)).all() and z[5, 3] == 0.0
    np.testing.assert_array_equal(z[:5, 3], 0.0)
    zz = np.clip(z[np.isfinite(z)], -50, 50)
    got = host(ops.double_sigmoid_abs(dev(zz)))
    np.testing.assert_allclose(got, np.abs(o_norm.double_sigmoid(zz)), rtol=1e-12, atol=1e-15)
    # even number of controls -> mean of the two middle values
    ctrl2 = np.zeros(W, bool)
    ctrl2[:4] = True
    np.testing.assert_allclose(host(ops.mad_robustize(dev(prof[:, :2].copy()), dev(ctrl2.astype(np.uint8)))),
                               o_norm.mad_robustize(prof[:, :2], ctrl2), rtol=1e-12)
