"""The oracle restatements against fixtures produced by the reference's own functions
(oracle/make_golden.py) -- this is what pins the oracle (task section 3)."""
import os

import numpy as np
import pytest

from oracle import cosine, lanczos, preprocess, qc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_maxproj_matches_reference(golden_dir):
    g = _load(golden_dir, "maxproj.npz")
    for k in ("z3", "z5", "z1"):
        out = preprocess.max_projection(list(g[f"{k}_in"]))
        assert out.dtype == np.uint16
        np.testing.assert_array_equal(out, g[f"{k}_out"])
        np.testing.assert_array_equal(preprocess.max_projection_field(g[f"{k}_in"][None])[0], g[f"{k}_out"])
    assert int(g["mismatch_raises"]) == 1
    with pytest.raises(ValueError):
        preprocess.max_projection([np.zeros((4, 4), np.uint16), np.zeros((4, 5), np.uint16)])


@pytest.mark.parametrize("case", ["a", "b", "c", "small", "mid"])
def test_illum_qc_matches_reference(golden_dir, case):
    g = _load(golden_dir, "illum_qc.npz")
    img, ill = g[f"{case}_img"], g[f"{case}_illum"]
    corr = preprocess.illum_correct(img, ill)
    assert corr.dtype == np.float64
    labels, mag, pw = qc.radial_power_spectrum(corr)
    np.testing.assert_array_equal(np.asarray(labels), g[f"{case}_rps_labels"])
    np.testing.assert_allclose(np.asarray(mag, float), g[f"{case}_rps_mag"], rtol=1e-10)
    np.testing.assert_allclose(np.asarray(pw, float), g[f"{case}_rps_pow"], rtol=1e-10)
    for tag, im in (("corr", corr), ("raw", img.astype(float))):
        s = qc.power_loglog_slope(im)
        ref = float(g[f"{case}_slope_{tag}"])
        if np.isnan(ref):
            assert np.isnan(s)          # min(H,W) < 24: the reference's latent NaN
        else:
            assert s == pytest.approx(ref, rel=1e-9, abs=1e-12)
        assert qc.percent_maximal(im) == float(g[f"{case}_pct_{tag}"])


def test_illum_shape_mismatch_is_ignored(golden_dir):
    g = _load(golden_dir, "illum_qc.npz")
    img = g["mismatch_img"]
    corr = preprocess.illum_correct(img, np.ones((8, 8)))
    np.testing.assert_array_equal(corr, img.astype(float))
    assert qc.percent_maximal(corr) == float(g["mismatch_pct"])
    assert qc.power_loglog_slope(corr) == pytest.approx(float(g["mismatch_slope"]), rel=1e-9)


def test_constant_image_known_answers(golden_dir):
    g = _load(golden_dir, "illum_qc.npz")
    assert float(g["const_slope"]) == 0.0 and float(g["const_pct"]) == 100.0
    const = np.full((48, 48), 1234.0)
    assert qc.power_loglog_slope(const) == 0.0
    assert qc.percent_maximal(const) == 100.0
    # 24 <= min(H,W) < 40 -> at most 2 rings -> slope 0.0; < 24 -> NaN
    rng = np.random.default_rng(0)
    assert qc.power_loglog_slope(rng.random((30, 36))) == 0.0
    assert np.isnan(qc.power_loglog_slope(rng.random((20, 30))))
    assert qc.ring_labels(2160, 2160).size == 268 and qc.ring_labels(1080, 1080).size == 133


@pytest.mark.parametrize("case", ["x2", "x4", "odd", "full"])
def test_lanczos_matches_reference(golden_dir, case):
    g = _load(golden_dir, "rebin_lanczos.npz")
    src, ref = g[f"{case}_in"], g[f"{case}_out"]
    np.testing.assert_array_equal(lanczos.pil_resize(src, ref.shape), ref)
    np.testing.assert_array_equal(lanczos.resize_restated(src, ref.shape), ref)


def test_lanczos_overflow_quirk_present(golden_dir):
    g = _load(golden_dir, "rebin_lanczos.npz")
    src, ref = g["full_in"], g["full_out"]
    # overshoot next to the saturated plateau is stored as 0xFF00 | low byte, not 65535
    hi = ref[(ref >= 0xFF00) & (ref != 65535)]
    assert hi.size > 0


@pytest.mark.parametrize("case", ["g4", "g7", "g2", "g1"])
def test_cosine_matches_sklearn(golden_dir, case):
    g = _load(golden_dir, "cosine.npz")
    x = g[f"{case}_x"]
    np.testing.assert_allclose(cosine.triu_values(x), g[f"{case}_triu"], rtol=1e-12, atol=1e-15)
    m = cosine.triu_mean(x)
    ref = float(g[f"{case}_mean"])
    if np.isnan(ref):
        assert np.isnan(m)
    else:
        assert m == pytest.approx(ref, rel=1e-12, abs=1e-15)
        n = x.shape[0]
        assert cosine.triu_sum_closed_form(x) * 2 / (n * (n - 1)) == pytest.approx(ref, rel=1e-10, abs=1e-14)


@pytest.mark.parametrize("case", ["a", "b", "const", "tiny"])
def test_scale_to_8bit_matches_reference(golden_dir, case):
    from oracle import crops
    g = _load(golden_dir, "crops.npz")
    out = crops.scale_to_8bit(g[f"{case}_in"])
    assert out.dtype == np.uint8
    np.testing.assert_array_equal(out, g[f"{case}_out"])


def test_well_mean_matches_reference_aggregation(golden_dir):
    """tests/golden/well_agg.npz: the per-well table written by the reference's own
    Normalize_CP_ami.concatenate_csv_from_s3 (pycytominer stubbed out, oracle/make_golden.py).  The
    oracle's well_mean on the object tables reproduces its columns (mean, all images kept), NaN cells
    skipped per column as pandas does."""
    import io
    import pandas as pd
    from oracle import normalize as o_norm
    g = np.load(os.path.join(golden_dir, "well_agg.npz"))
    out = pd.read_csv(io.BytesIO(g["out_mean_all"].tobytes()))
    image = pd.read_csv(io.BytesIO(g["in_Image"].tobytes()))
    for name, prefix in (("Nuclei", "DNA_"), ("Cells", "Cell_"), ("Cytoplasm", "Cyto_")):
        tab = pd.read_csv(io.BytesIO(g[f"in_{name}"].tobytes())).merge(image[["ImageNumber", "Metadata_Well"]], on="ImageNumber")
        feats = [c for c in tab.columns if c not in ("ImageNumber", "Metadata_Well")]
        assert tab[feats].isna().to_numpy().any()                # the fixture does exercise NaN skipping
        wells, means = o_norm.well_mean(tab[feats].to_numpy(float), tab["Metadata_Well"].to_numpy())
        ref = out.set_index("Metadata_Well").loc[wells, [prefix + f for f in feats]].to_numpy(float)
        np.testing.assert_allclose(means, ref, rtol=1e-12)
    assert str(g["method"]) == "mad_robustize"
    assert str(g["samples_query"]) == "Metadata_Compound == 'DMSO' and Metadata_Timepoint == '24h'"


@pytest.mark.parametrize("tag", ["global", "pertime"])
def test_double_sigmoid_and_group_cosine_match_reference_script(golden_dir, tag):
    """tests/golden/cosine_script.npz: the files written by the reference's own
    Feature_select_cosine_ami.concatenate_normalized_csv_from_s3 with an identity feature selection
    (oracle/make_golden.py).  The oracle's double sigmoid and per-group mean cosine reproduce them."""
    import io
    import pandas as pd
    from oracle import normalize as o_norm
    g = np.load(os.path.join(golden_dir, "cosine_script.npz"))
    raw = pd.read_csv(io.BytesIO(g[f"{tag}_EXP_CP_features_selected_allTimes_raw.csv"].tobytes()))
    dsig = pd.read_csv(io.BytesIO(g[f"{tag}_EXP_CP_features_selected_allTimes_dSig.csv"].tobytes()))
    avg = pd.read_csv(io.BytesIO(g[f"{tag}_EXP_Average_cosine_similarity.csv"].tobytes()))
    feats = [c for c in raw.columns if "Metadata" not in c]
    # pandas writes small floats with 15 decimals: the file holds the values to 1e-15 absolute
    np.testing.assert_allclose(np.abs(o_norm.double_sigmoid(raw[feats].to_numpy(float))), dsig[feats].to_numpy(float),
                               rtol=1e-13, atol=2e-15, equal_nan=True)
    keys = ["Metadata_Compound", "Metadata_Timepoint", "Metadata_ConcLevel"]
    assert len(avg) == len(dsig[keys].drop_duplicates())
    for _, row in avg.iterrows():
        grp = dsig[(dsig[keys[0]] == row[keys[0]]) & (dsig[keys[1]] == row[keys[1]]) & (dsig[keys[2]] == row[keys[2]])]
        want = cosine.triu_mean(grp[feats].fillna(0).to_numpy(float))
        np.testing.assert_allclose(want, row["average_cosine_similarity"], rtol=1e-12, equal_nan=True)
