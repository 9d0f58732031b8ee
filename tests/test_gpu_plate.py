"""Dense row packing and well aggregation on the device vs NumPy / pandas (single GPU)."""
import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import normalize as o_norm
from oracle import object_stats as o_obj
from oracle import preprocess as o_pre
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu


def test_pack_rows_and_well_means_match_groupby():
    torch = require_gpu()
    from image_processing_suite_b200 import ops, plate
    F, C, H, W, cells = 6, 2, 96, 128, 9
    labs = np.stack([synth.make_labels(H, W, cells if f != 2 else 0, seed=f, amin=5, amax=10) for f in range(F)])
    mps = np.stack([o_pre.max_projection_field(synth.field_numpy(labs[f], c=C, z=2, seed=f)) for f in range(F)])
    res = ops.object_stats(dev(labs), dev(mps), None, 1.0, n_max=cells + 2)
    field_well = np.array([4, 4, 7, 7, 9, 9], np.int32)
    rows, total = plate.pack_rows(res["ints"], res["flts"], res["n_objects"], dev(field_well), field_base=100)
    n = int(total.item())
    expect = []
    for f in range(F):
        e_i, e_f = o_obj.object_stats(labs[f], mps[f], None, 1.0)
        for r in range(e_i.shape[0]):
            expect.append(np.r_[field_well[f], 100 + f, e_i[r], e_f[r]])
    expect = np.asarray(expect)
    assert n == expect.shape[0] and rows.shape[1] == 10 + 5 * C
    got = host(rows)[:n].astype(np.float64)
    np.testing.assert_array_equal(got[:, :8], expect[:, :8])           # well, field, label, area, bbox exact
    np.testing.assert_allclose(got[:, 8:], expect[:, 8:], rtol=1e-5, atol=1e-6)
    # one-rank "gather" + per-well means == pandas groupby over the same rows
    g = plate.RowGatherer(rows.shape[0], rows.shape[1])
    all_rows, counts = g.gather(rows, n)
    mean, count = plate.well_means(all_rows, counts, 12)
    ids, ref = o_norm.well_mean(got, got[:, 0].astype(int))
    m, c = host(mean), host(count)
    np.testing.assert_allclose(m[ids], ref, rtol=1e-12)
    assert c.sum() == n and set(np.flatnonzero(c)) == set(ids.tolist())
    assert np.isnan(m[0]).all()


def test_well_aggregator_streaming_equals_one_shot():
    torch = require_gpu()
    from image_processing_suite_b200 import plate
    rng = np.random.default_rng(3)
    blocks, cap, D, n_wells = 5, 40, 12, 9
    rows = rng.normal(size=(blocks, cap, D)).astype(np.float32)
    rows[:, :, 0] = rng.integers(0, n_wells, (blocks, cap))
    counts = np.array([40, 0, 17, 33, 1], np.int64)
    d_rows, d_counts = dev(rows), dev(counts)
    one_mean, one_count = plate.well_means(d_rows, d_counts, n_wells)
    agg = plate.WellAggregator(n_wells, D)
    agg.add(d_rows[:2].contiguous(), d_counts[:2].contiguous())
    agg.add(d_rows[2:].contiguous(), d_counts[2:].contiguous())
    mean, count = agg.finalize()
    np.testing.assert_allclose(host(mean), host(one_mean), rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(host(count), host(one_count))
    valid = np.concatenate([rows[b, :counts[b]] for b in range(blocks)]).astype(np.float64)
    ids, ref = o_norm.well_mean(valid, valid[:, 0].astype(int))
    np.testing.assert_allclose(host(mean)[ids], ref, rtol=1e-12)
