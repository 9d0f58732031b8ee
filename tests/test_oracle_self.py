"""Oracle self-consistency for the unpinned rows (A3', A4, A5, normalisation)."""
import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import illum, normalize, object_stats, preprocess


def test_sum_bin_exact():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 65536, (3, 24, 32), dtype=np.uint16)
    a[0, :4, :4] = 65535
    for b in (1, 2, 4):
        s = preprocess.sum_bin(a, b)
        assert s.dtype == np.uint32 and s.shape == (3, 24 // b, 32 // b)
        assert int(s.sum()) == int(a.astype(np.uint64).sum())
    assert preprocess.sum_bin(a, 4)[0, 0, 0] == 16 * 65535
    with pytest.raises(ValueError):
        preprocess.sum_bin(a[:, :23], 2)


def test_object_stats_two_restatements_agree():
    lab = synth.make_labels(160, 200, 25, seed=3)
    raw = synth.field_numpy(lab, c=3, z=2, seed=3)
    ill = synth.make_illum(3, 160, 200, seed=3)
    mp = preprocess.max_projection_field(raw)
    for il, sc in ((None, 1.0), (ill, 1.0 / 65535.0)):
        i1, f1 = object_stats.object_stats(lab, mp, il, sc)
        i2, f2 = object_stats.object_stats_bincount(lab, mp, il, sc)
        np.testing.assert_array_equal(i1, i2)
        np.testing.assert_allclose(f1, f2, rtol=1e-9, atol=1e-12)
    assert i1.shape[0] == lab.max()


def test_object_stats_conventions():
    lab = np.zeros((8, 10), np.int32)
    lab[2:5, 3:7] = 2            # label 1 absent -> skipped
    lab[6, 0] = 5
    mp = np.arange(80, dtype=np.uint16).reshape(1, 8, 10)
    ints, flts = object_stats.object_stats(lab, mp)
    np.testing.assert_array_equal(ints, [[2, 12, 2, 3, 5, 7], [5, 1, 6, 0, 7, 1]])
    v = mp[0, 2:5, 3:7].astype(float)
    np.testing.assert_allclose(flts[0], [3.0, 4.5, v.sum(), v.mean(), v.std(), v.min(), v.max()])
    np.testing.assert_allclose(flts[1], [6.0, 0.0, 60, 60, 0, 60, 60])
    e_i, e_f = object_stats.object_stats(np.zeros((4, 4), np.int32), mp[:, :4, :4])
    assert e_i.shape == (0, 6) and e_f.shape == (0, 7)


def test_illum_estimation_recovers_vignette():
    lab = np.zeros((96, 128), np.int32)
    fields = np.stack([preprocess.max_projection_field(synth.field_numpy(lab, c=2, z=1, seed=s, saturate_frac=0))
                       for s in range(12)])
    est = illum.estimate(fields, sigma=6.0)
    assert est.shape == (2, 96, 128) and est.min() >= 1.0
    # brighter in the centre than at the corners (the generator's vignette)
    assert est[0, 48, 64] > est[0, 2, 2]
    acc = illum.accumulate(fields)
    np.testing.assert_allclose(illum.estimate_from_sum(acc, 12, 6.0), est, rtol=1e-12)
    med = illum.estimate(fields, sigma=6.0, mode="median")
    assert np.abs(med / est - 1).max() < 0.05


def test_normalize_restatements():
    rng = np.random.default_rng(5)
    rows = rng.normal(size=(200, 6))
    wells = rng.integers(0, 7, 200)
    ids, means = normalize.well_mean(rows, wells)
    for k, wid in enumerate(ids):
        np.testing.assert_allclose(means[k], rows[wells == wid].mean(axis=0), rtol=1e-12)
    ctrl = np.zeros(7, bool)
    ctrl[:3] = True
    z = normalize.mad_robustize(means, ctrl)
    np.testing.assert_allclose(np.median(z[:3], axis=0), 0, atol=1e-12)
    ds = normalize.double_sigmoid(np.array([-10.0, 0.0, 2.3538, 10.0]))
    np.testing.assert_allclose(ds[[1, 2]], [0.0, 1 / np.sqrt(2)])
    assert abs(ds[0]) < 1 and abs(ds[3]) < 1
