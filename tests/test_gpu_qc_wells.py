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


def test_well_mean_wide_table_and_nan_skipping_match_pandas():
    """CellProfiler tables carry > 255 columns and NaN cells; pandas skips NaN per
    column (Normalize_CP_ami.py:126).  Float32 and float64 entry points."""
    require_gpu()
    import pandas as pd
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(11)
    n_wells, D, N = 17, 1031, 3000
    wells = rng.integers(0, n_wells, N).astype(np.int32)
    rows = (rng.normal(3.0, 2.0, (N, D)) * 10.0 ** rng.integers(-6, 7, (1, D))).astype(np.float32)
    rows[rng.random((N, D)) < 0.02] = np.nan
    rows[:, 5] = np.nan                                         # an all-NaN column
    rows[wells == 3, 9] = np.nan                                # all-NaN for one well only
    rows[7, 11] = np.inf
    ref = pd.DataFrame(rows.astype(np.float64)).groupby(wells).mean()
    mean, count = ops.well_mean(dev(rows), dev(wells), n_wells)
    m = host(mean)
    np.testing.assert_allclose(m[ref.index.to_numpy()], ref.to_numpy(), rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(host(count), np.bincount(wells, minlength=n_wells))
    assert np.isnan(m[:, 5]).all() and np.isnan(m[3, 9]) and np.isinf(m[wells[7], 11])
    # exact accumulation: any row order gives the same bits
    perm = rng.permutation(N)
    mean2, _ = ops.well_mean(dev(rows[perm]), dev(wells[perm]), n_wells)
    np.testing.assert_array_equal(host(mean2).view(np.int64), m.view(np.int64))
    # float64 rows
    rows64 = rows.astype(np.float64) * (1.0 + 1e-9 * rng.random((N, D)))
    ref64 = pd.DataFrame(rows64).groupby(wells).mean()
    mean64, _ = ops.well_mean_f64(dev(rows64), dev(wells), n_wells)
    np.testing.assert_allclose(host(mean64)[ref64.index.to_numpy()], ref64.to_numpy(), rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("D", [3, 40, 77])
def test_well_median_matches_pandas(D):
    require_gpu()
    import pandas as pd
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(12 + D)
    n_wells, N = 13, 4001
    wells = rng.integers(0, n_wells, N).astype(np.int32)
    wells[wells == 4] = 5                                       # empty well
    rows = rng.normal(0.0, 50.0, (N, D))
    rows[rng.random((N, D)) < 0.05] = np.nan
    rows[:, 1] = np.round(rows[:, 1])                           # ties
    rows[wells == 2, 0] = np.nan
    rows[::7, 2] = -0.0
    ref = pd.DataFrame(rows).groupby(wells).median()
    med, count = ops.well_median_f64(dev(rows), dev(wells), n_wells)
    m = host(med)
    np.testing.assert_array_equal(m[ref.index.to_numpy()], ref.to_numpy())
    assert np.isnan(m[4]).all() and host(count)[4] == 0
    np.testing.assert_array_equal(host(count), np.bincount(wells, minlength=n_wells))


def test_header_blocks_pack_count_and_aggregate():
    """ips_pack_rows_block -> ips_block_counts -> ips_well_sums_add_blocks: the no-host-sync form
    of pack_rows + well_means gives the same per-well means, bit for bit."""
    torch = require_gpu()
    from image_processing_suite_b200 import plate
    rng = np.random.default_rng(21)
    F, n_max, C = 7, 30, 3
    nf = 2 + 5 * C
    D = 8 + nf
    n_obj = np.array([30, 0, 12, -1, 5, 30, 1], np.int32)
    ints = rng.integers(0, 500, (F, n_max, 6)).astype(np.int32)
    flts = rng.normal(10.0, 3.0, (F, n_max, nf)).astype(np.float32)
    field_well = np.array([2, 2, 3, 3, 3, 8, 8], np.int32)
    d = [dev(x) for x in (ints, flts, n_obj, field_well)]
    rows, total = plate.pack_rows(*d, field_base=40)
    n = int(total.item())
    table = torch.zeros((3, F * n_max + 1, D), dtype=torch.float32, device="cuda")
    table[2].fill_(float("nan"))                                # block 2 stays header-less garbage with count 0
    table[2, 0].zero_()
    plate.pack_rows_block(*d, table[1], field_base=40)
    counts = host(plate.block_counts(table))
    assert counts.tolist() == [0, n, 0]
    assert torch.equal(table[1, 1:n + 1], rows[:n])
    agg = plate.WellAggregator(10, D)
    agg.add_blocks(table)
    mean_b, count_b = agg.finalize()
    ids = torch.full((F * n_max,), -1, dtype=torch.int32, device="cuda")
    ids[:n] = rows[:n, 0].to(torch.int32)
    from image_processing_suite_b200 import ops
    mean_r, count_r = ops.well_mean(rows, ids, 10)
    np.testing.assert_array_equal(host(mean_b).view(np.int64), host(mean_r).view(np.int64))
    np.testing.assert_array_equal(host(count_b), host(count_r))


@pytest.mark.parametrize("hw", [(64, 64), (96, 81), (45, 130), (216, 216), (250, 333)])
@pytest.mark.parametrize("kind", ["float64", "uint16", "uint16/illum"])
def test_rps_on_device_matches_reference_arithmetic(hw, kind):
    """ops.rps_spectrum + ops.loglog_slope (exact radix-select median, real FFT, ring sums over the
    Hermitian half, device least squares) vs the oracle's rps / slope (Illumination_QC_mult.py:31-116)."""
    require_gpu()
    import torch
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(hash((hw, kind)) % 2 ** 31)
    H, W = hw
    yy, xx = np.mgrid[0:H, 0:W]
    base = 2000.0 + 900.0 * np.sin(yy / 7.0) * np.cos(xx / 5.0) + rng.normal(0, 60.0, hw)
    if kind == "float64":
        img, ill, x = base, None, base
        mag, pw = ops.rps_spectrum(dev(img))
    else:
        img = np.clip(np.rint(base), 0, 65535).astype(np.uint16)
        ill = (1.0 + 0.4 * rng.random(hw)) if kind.endswith("illum") else None
        x = img.astype(np.float64) / ill if ill is not None else img.astype(np.float64)
        mag, pw, corr = ops.rps_spectrum(dev(img), dev(ill) if ill is not None else None, want_corrected=True)
        np.testing.assert_array_equal(host(corr), x)                      # the float64 divide of :145-150, exact
    labels, e_mag, e_pow = o_qc.radial_power_spectrum(x)
    np.testing.assert_allclose(host(mag), e_mag, rtol=1e-9)
    np.testing.assert_allclose(host(pw), e_pow, rtol=1e-9)
    slope = float(ops.loglog_slope(pw).item())
    assert slope == pytest.approx(o_qc.power_loglog_slope(x), rel=1e-9, abs=1e-12)


def test_rps_on_device_degenerate_cases():
    require_gpu()
    import torch
    from image_processing_suite_b200 import ops
    # constant image: no normalisation, zero spectrum, slope 0.0
    mag, pw = ops.rps_spectrum(dev(np.full((48, 64), 7.0)))
    assert float(pw.abs().max()) == 0.0 and float(ops.loglog_slope(pw).item()) == 0.0
    # 24 <= min(H, W) < 40: at most two rings -> 0.0
    rng = np.random.default_rng(1)
    mag, pw = ops.rps_spectrum(dev(rng.random((33, 50))))
    assert pw.numel() == 2 and float(ops.loglog_slope(pw).item()) == 0.0
    # min(H, W) < 24: no ring
    mag, pw = ops.rps_spectrum(dev(rng.random((20, 30))))
    assert pw.numel() == 0
    # median of an even count with the two middle values apart, and with heavy ties
    x = np.zeros((40, 40))
    x[:20] = 10.0
    x[0, 0] = 500.0
    labels, e_mag, e_pow = o_qc.radial_power_spectrum(x)
    mag, pw = ops.rps_spectrum(dev(x))
    np.testing.assert_allclose(host(pw), e_pow, rtol=1e-9, atol=1e-9)
