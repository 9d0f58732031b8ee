"""K2 parity: illumination-function estimation (CUDA) vs oracle/illum.py (float64 SciPy).

Integer sums bit-exact; the function itself within RTOL = 1e-5 relative (north_star).
PARITY UNPINNED against the reference (it holds no estimation code).
"""
import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import illum as o_illum
from oracle import preprocess as o_pre
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _plate(F, C, H, W, seed=0):
    lab = np.zeros((H, W), np.int32)
    return np.stack([o_pre.max_projection_field(synth.field_numpy(lab, c=C, z=2, seed=seed + s, saturate_frac=1e-3))
                     for s in range(F)])


@pytest.mark.parametrize("shape,sigma", [((13, 2, 96, 128), 6.0), ((5, 3, 70, 50), 2.5), ((4, 1, 64, 200), 21.3)])
def test_illum_mean_mode(shape, sigma):
    require_gpu()
    from image_processing_suite_b200 import ops
    F, C, H, W = shape
    fields = _plate(F, C, H, W)
    est = ops.IllumEstimator(C, H, W)
    est.add(dev(fields[:3]))
    est.add(dev(fields[3:]))
    np.testing.assert_array_equal(host(est.acc).astype(np.uint64), o_illum.accumulate(fields))
    got = host(est.finalize(sigma, 0.02))
    ref = o_illum.estimate(fields, sigma, 0.02)
    assert got.min() >= 1.0
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=0)


def test_illum_ragged_and_robust_fracs():
    require_gpu()
    from image_processing_suite_b200 import ops
    fields = _plate(6, 2, 37, 53, seed=3)             # C*H*W % 8 != 0 -> scalar accumulate
    est = ops.IllumEstimator(2, 37, 53).add(dev(fields))
    np.testing.assert_array_equal(host(est.acc).astype(np.uint64), o_illum.accumulate(fields))
    for frac in (0.0, 0.02, 0.5, 0.999):
        np.testing.assert_allclose(host(est.finalize(3.0, frac)), o_illum.estimate(fields, 3.0, frac), rtol=RTOL)


@pytest.mark.parametrize("N", [1, 2, 7, 12])
def test_illum_median_mode(N):
    require_gpu()
    from image_processing_suite_b200 import ops
    fields = _plate(N, 2, 48, 64, seed=9)
    raw = host(ops.illum_median(dev(fields)))
    np.testing.assert_array_equal(raw.astype(np.float64), np.median(fields, axis=0))
    got = host(ops.illum_smooth_rescale(dev(raw), 4.0, 0.02))
    np.testing.assert_allclose(got, o_illum.estimate(fields, 4.0, 0.02, mode="median"), rtol=RTOL)


def test_illum_all_zero_plate_gives_ones():
    require_gpu()
    from image_processing_suite_b200 import ops
    est = ops.IllumEstimator(1, 32, 32).add(dev(np.zeros((3, 1, 32, 32), np.uint16)))
    np.testing.assert_array_equal(host(est.finalize(2.0)), np.ones((1, 32, 32), np.float32))


def test_illum_linearity_at_full_size():
    """Config-3 size: accumulate is linear -- sum over two batches == sum over their union,
    and equals torch's own int64 sum (a checksum of checksums)."""
    torch = require_gpu()
    from image_processing_suite_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    fields = torch.randint(0, 65536, (6, 5, 2160, 2160), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
    a = ops.IllumEstimator(5, 2160, 2160).add(fields[:2]).add(fields[2:])
    b = ops.IllumEstimator(5, 2160, 2160).add(fields)
    assert bool((a.acc.view(torch.int32) == b.acc.view(torch.int32)).all())
    assert bool((b.acc.view(torch.int32).to(torch.int64) == fields.to(torch.int64).sum(dim=0)).all())
    out = b.finalize(sigma=40.0)
    assert float(out.min()) >= 1.0 and bool(torch.isfinite(out).all())
