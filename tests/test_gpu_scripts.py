"""Drop-in scripts end to end on a local directory (IPS_STORAGE_ROOT), CUDA path vs the
oracle and vs the golden fixtures made by the reference's own functions."""
import io
import os

import numpy as np
import pandas as pd
import pytest

from image_processing_suite_b200 import synth
from oracle import illum as o_illum
from oracle import lanczos as o_lz
from oracle import object_stats as o_obj
from oracle import preprocess as o_pre
from oracle import qc as o_qc
from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu


def test_max_projection_script(tmp_path, monkeypatch):
    require_gpu()
    from image_processing_suite_b200.scripts import MaxProjection, storage, tiffio
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    C, Z, H, W = 2, 3, 40, 56
    rng = np.random.default_rng(1)
    rows, stacks = [], {}
    for field in range(2):
        st = rng.integers(0, 65536, (C, Z, H, W), dtype=np.uint16)
        stacks[field] = st
        for p in range(Z):                                   # plane-major, channel-minor rows
            for j in range(C):
                name = f"f{field}p{p}c{j}.tiff"
                # per channel group: raw strips (with one LZW plane mixed in), LZW (device codec), deflate (host decode)
                comp = (None, "tiff_lzw", "tiff_adobe_deflate")[(field + j) % 3]
                if comp is None and p == 1:
                    comp = "tiff_lzw"
                s3.upload_fileobj(io.BytesIO(tiffio.encode(st[j, p], comp)), "img", f"exp/Images/{name}")
                rows.append({"PlateID": "P1", "Image_PathName": "exp/Images", "Image_FileName": name})
    rows.append({"PlateID": "P1", "Image_PathName": "exp/Images", "Image_FileName": "tail.tiff"})   # incomplete chunk
    csv_bytes = pd.DataFrame(rows).to_csv(index=False, sep=";").encode()
    s3.upload_fileobj(io.BytesIO(csv_bytes), "sets", "d.csv")
    n = MaxProjection.run("sets", "d.csv", C, Z, "img", s3)
    assert n == 4
    for field in range(2):
        for j in range(C):
            out = tiffio.decode(s3.get_object(Bucket="img", Key=f"exp/ImagesStacked/f{field}p0c{j}.tiff")["Body"].read())
            np.testing.assert_array_equal(out, o_pre.max_projection(list(stacks[field][j])))
    # the per-group function with the reference's error contract
    with pytest.raises(FileNotFoundError):
        MaxProjection.max_projection(["exp/Images/missing.tiff"], "img", s3)
    s3.upload_fileobj(io.BytesIO(tiffio.encode(np.zeros((8, 9), np.uint16))), "img", "exp/Images/odd.tiff")
    with pytest.raises(ValueError, match="shape mismatch"):
        MaxProjection.max_projection(["exp/Images/f0p0c0.tiff", "exp/Images/odd.tiff"], "img", s3)


def test_rebinning_script(tmp_path, monkeypatch, golden_dir):
    require_gpu()
    from image_processing_suite_b200.scripts import Image_rebinning, storage, tiffio
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    g = np.load(os.path.join(golden_dir, "rebin_lanczos.npz"))
    src = g["x2_in"]
    out = tiffio.decode(Image_rebinning.process_image_in_memory(tiffio.encode(src), (src.shape[1] // 2, src.shape[0] // 2)))
    np.testing.assert_array_equal(out, g["x2_out"])
    s3 = storage.client()
    rng = np.random.default_rng(2)
    img = rng.integers(0, 65536, (120, 120), dtype=np.uint16)
    s3.put_object(Bucket="b", Key="e/Image/a.tiff", Body=tiffio.encode(img))
    s3.put_object(Bucket="b", Key="e/Image/notes.txt", Body=b"skip me")
    s3.put_object(Bucket="b", Key="e/Image/bad.tiff", Body=b"not a tiff")
    assert Image_rebinning.process_images_in_s3("b", "e/Image", 60) == 1      # bad file logged and skipped
    got = tiffio.decode(s3.get_object(Bucket="b", Key="e/Image_binned/a.tiff")["Body"].read())
    np.testing.assert_array_equal(got, o_lz.pil_resize(img, (60, 60)))


def test_illumination_qc_script(tmp_path, golden_dir):
    require_gpu()
    from image_processing_suite_b200.scripts import Illumination_QC_mult as qc, tiffio
    g = np.load(os.path.join(golden_dir, "illum_qc.npz"))
    # reference-function goldens: slope and PercentMaximal of corrected images
    os.makedirs(tmp_path / "img")
    os.makedirs(tmp_path / "illum")
    rows = []
    for case in ("a", "b", "c"):
        tiffio.write(str(tmp_path / "img" / f"{case}.tiff"), g[f"{case}_img"])
        rows.append({"FileName_CH1": f"{case}.tiff", "Metadata_Well": case})
    rows.append({"FileName_CH1": "missing.tiff", "Metadata_Well": "z"})
    # one illumination function per channel: use case a's; b and c have other shapes or the same
    np.save(tmp_path / "illum" / "CH1_illum.npy", g["a_illum"])
    pd.DataFrame(rows).to_csv(tmp_path / "load.csv", index=False)
    out = qc.main(["--load-data", str(tmp_path / "load.csv"), "--data-path", str(tmp_path / "img"),
                   "--illum-path", str(tmp_path / "illum"), "--channels", "CH1",
                   "--output", str(tmp_path / "qc.csv"), "--threads", "3"])
    assert out.loc[3, "QC_Error_CH1"] == "File Not Found"
    ref_a_slope, ref_a_pct = float(g["a_slope_corr"]), float(g["a_pct_corr"])
    assert out.loc[0, "ImageQuality_PowerLogLogSlope_CH1"] == pytest.approx(ref_a_slope, rel=1e-6)
    assert out.loc[0, "ImageQuality_PercentMaximal_CH1"] == ref_a_pct
    for k, case in ((1, "b"), (2, "c")):
        img, ill = g[f"{case}_img"], g["a_illum"]
        corr = o_pre.illum_correct(img, ill)                 # shape mismatch -> uncorrected
        assert out.loc[k, "ImageQuality_PowerLogLogSlope_CH1"] == pytest.approx(o_qc.power_loglog_slope(corr), rel=1e-6)
        assert out.loc[k, "ImageQuality_PercentMaximal_CH1"] == o_qc.percent_maximal(corr)
    # degenerate cases of the reference
    assert qc.calculate_qc_metrics(np.full((48, 48), 7.0), "X") == {
        "ImageQuality_PowerLogLogSlope_X": 0.0, "ImageQuality_PercentMaximal_X": 100.0}
    assert np.isnan(qc.calculate_qc_metrics(np.random.default_rng(0).random((20, 30)), "X")["ImageQuality_PowerLogLogSlope_X"])
    saved = pd.read_csv(tmp_path / "qc.csv")
    assert "ImageQuality_PercentMaximal_CH1" in saved.columns and len(saved) == 4


def test_feature_extraction_and_illum_estimate_scripts(tmp_path):
    require_gpu()
    from image_processing_suite_b200.scripts import Feature_extraction as fe, Illumination_estimate as ie, tiffio
    C, H, W, cells, sites = 2, 96, 128, 10, 5
    chans = ["DNA", "ER"]
    os.makedirs(tmp_path / "img")
    rows, labs, mps = [], [], []
    for s in range(sites):
        lab = synth.make_labels(H, W, cells, seed=s, amin=5, amax=10)
        mp = o_pre.max_projection_field(synth.field_numpy(lab, c=C, z=2, seed=s))
        labs.append(lab); mps.append(mp)
        row = {"Metadata_Well": f"A{s // 2 + 1:02d}", "Metadata_Site": s % 2 + 1}
        for j, ch in enumerate(chans):
            tiffio.write(str(tmp_path / "img" / f"s{s}_{ch}.tiff"), mp[j])
            row[f"FileName_{ch}"] = f"s{s}_{ch}.tiff"
            row[f"PathName_{ch}"] = str(tmp_path / "img")
        tiffio.write(str(tmp_path / "img" / f"s{s}_mask.tiff"), lab.astype(np.uint16))
        row["Objects_FileName_Nuclei"] = f"s{s}_mask.tiff"
        row["Objects_PathName_Nuclei"] = str(tmp_path / "img")
        rows.append(row)
    pd.DataFrame(rows).to_csv(tmp_path / "load.csv", index=False)
    # 1. estimate illumination functions for the "plate"
    funcs = ie.main(["--load-data", str(tmp_path / "load.csv"), "--data-path", str(tmp_path / "img"),
                     "--illum-path", str(tmp_path / "illum"), "--channels", *chans, "--filter-size", "20", "--batch", "2"])
    ref = o_illum.estimate(np.stack(mps), 20 / 2.35, 0.02)
    for j, ch in enumerate(chans):
        np.testing.assert_allclose(funcs[ch], ref[j], rtol=1e-5)
        assert np.load(tmp_path / "illum" / f"{ch}_illum.npy").dtype == np.float32
    # 2. add them to the LoadData CSV (load_data_*_illum.csv) and run the CellProfiler-style CLI
    df = pd.read_csv(tmp_path / "load.csv")
    for ch in chans:
        df[f"FileName_Illum{ch}"] = f"{ch}_illum.npy"
        df[f"PathName_Illum{ch}"] = str(tmp_path / "illum")
    df.to_csv(tmp_path / "load_illum.csv", index=False)
    fe.main(["-c", "-r", "-p", "Feature_Extraction_CL2.0.cppipe", "-o", str(tmp_path / "out"),
             "--data-file", str(tmp_path / "load_illum.csv"), "--batch", "2"])
    image = pd.read_csv(tmp_path / "out" / "Image.csv")
    nuclei = pd.read_csv(tmp_path / "out" / "Nuclei.csv")
    assert list(image.ImageNumber) == [1, 2, 3, 4, 5] and "Metadata_Well" in image.columns
    assert str(nuclei.AreaShape_Area.dtype).startswith("int") and str(image.Count_Nuclei.dtype).startswith("int")
    ill = np.stack([funcs[ch] for ch in chans])
    for s in range(sites):
        e_i, e_f = o_obj.object_stats(labs[s], mps[s], ill, 1 / 65535.0)
        sub = nuclei[nuclei.ImageNumber == s + 1]
        assert int(image.Count_Nuclei[s]) == e_i.shape[0] == len(sub)
        np.testing.assert_array_equal(sub.ObjectNumber, e_i[:, 0])
        np.testing.assert_array_equal(sub.AreaShape_Area, e_i[:, 1])
        np.testing.assert_array_equal(sub.AreaShape_BoundingBoxMinimum_Y, e_i[:, 2])
        np.testing.assert_array_equal(sub.AreaShape_BoundingBoxMaximum_X, e_i[:, 5])
        np.testing.assert_allclose(sub.Location_Center_X, e_f[:, 1], rtol=1e-5)
        np.testing.assert_allclose(sub.Intensity_MeanIntensity_ER, e_f[:, 2 + 5 + 1], rtol=1e-5)
        np.testing.assert_allclose(sub.Intensity_StdIntensity_DNA, e_f[:, 2 + 2], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(sub.Intensity_IntegratedIntensity_DNA, e_f[:, 2 + 0], rtol=1e-5)


@pytest.mark.parametrize("qc_drop", [False, True])
def test_normalize_script_matches_pandas_and_oracle(tmp_path, monkeypatch, qc_drop):
    """Normalize_CP_ami drop-in on a local 'bucket': per-well means and robust-z scores against a
    pure pandas / oracle restatement of the same steps (pycytominer parts: parity unpinned)."""
    require_gpu()
    from functools import reduce
    from image_processing_suite_b200.scripts import Normalize_CP_ami as nz, storage
    from oracle import normalize as o_norm
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    rng = np.random.default_rng(7)
    wells = [f"{r}{c:02d}" for r in "ABCD" for c in range(1, 7)]
    rows_img, tabs = [], {"Nuclei": [], "Cells": [], "Cytoplasm": []}
    image_number = 0
    for w in wells:
        for site in range(1, 4 if w != "B03" else 3):          # one well has fewer sites
            image_number += 1
            bad = image_number % 11 == 0
            rows_img.append({"ImageNumber": image_number, "Metadata_Well": w, "Metadata_Site": site, "Metadata_Plate": "P1",
                             "Count_Nuclei": int(rng.integers(5, 30)), "ImageQC_Blurry": int(bad),
                             "ExecutionTime_X": 1.5, "Intensity_Mean": rng.normal(0.3, 0.05)})
            for name in tabs:
                for obj in range(int(rng.integers(3, 7))):
                    tabs[name].append({"ImageNumber": image_number, "ObjectNumber": obj + 1,
                                       "AreaShape_Area": int(rng.integers(300, 1500)),
                                       "Intensity_MeanIntensity_DNA": rng.normal(0.2, 0.03),
                                       "Location_Center_X": rng.uniform(0, 2160)})
    s3.put_object(Bucket="b", Key="exp/P1/24h/Image.csv", Body=pd.DataFrame(rows_img).to_csv(index=False).encode())
    for name, rows in tabs.items():
        s3.put_object(Bucket="b", Key=f"exp/P1/24h/{name}.csv", Body=pd.DataFrame(rows).to_csv(index=False).encode())
    pm = pd.DataFrame({"Metadata_Well": wells, "Metadata_Plate": "P1", "Metadata_ConcLevel": 1,
                       "Metadata_Compound": ["dmso" if i % 4 == 0 else f"cmp{i}" for i in range(len(wells))]})
    s3.put_object(Bucket="b", Key="exp/Plate_P1_PlateMap.csv", Body=pm.to_csv(index=False).encode())
    for agg in ("mean", "median"):                         # --well_agg_func (Normalize_CP_ami.py:126,163)
        keys = nz.concatenate_csv_from_s3("b", ["P1"], ["24h"], "exp", "out", "DMSO", "norm", agg, False, qc_drop, s3)
        assert keys == ["norm/P1/Normalized_features_24h.csv"]
        got = pd.read_csv(io.BytesIO(s3.get_object(Bucket="out", Key=keys[0])["Body"].read()))
        # ---- the same steps in pandas + the oracle's mad_robustize -------------------------------
        image_df = pd.DataFrame(rows_img)
        failing = image_df.loc[image_df.filter(like="ImageQC_").any(axis=1), "ImageNumber"]
        per_table = []
        for name, prefix in nz.TABLE_PREFIX.items():
            df = image_df if name == "Image" else pd.DataFrame(tabs[name]).merge(
                image_df[["ImageNumber", "Metadata_Well", "Metadata_Site"]], on="ImageNumber", how="left")
            if qc_drop:
                df = df[~df["ImageNumber"].isin(failing)]
            keep = {"Metadata_Well", "Metadata_Site"} if qc_drop else {"Metadata_Well"}
            df = df.drop(columns=[c for c in df.columns if c == "ImageNumber" or (c.startswith("Metadata") and c not in keep)
                                  or any(s in c for s in nz.DROP_SUBSTRINGS)])
            df = df.rename(columns=lambda x: prefix + x if not x.startswith("Metadata_") else x)
            if qc_drop:
                sc = df.groupby("Metadata_Well")["Metadata_Site"].nunique()
                df = df.merge((sc.max() / sc).rename("scaling_factor"), on="Metadata_Well")
                ints = [c for c in df.select_dtypes(include="integer").columns if not c.startswith("Metadata")]
                df[ints] = df[ints].multiply(df["scaling_factor"], axis=0)
                df = df.drop(columns=["scaling_factor", "Metadata_Site"])
            per_table.append(df.groupby("Metadata_Well", as_index=False).agg(agg))
        merged = reduce(lambda l, r: pd.merge(l, r, on="Metadata_Well", how="outer"), per_table)
        pm2 = pm[["Metadata_Compound", "Metadata_ConcLevel", "Metadata_Well", "Metadata_Plate"]].copy()
        pm2["Metadata_Compound"] = pm2["Metadata_Compound"].str.upper()
        merged = pm2.merge(merged, on="Metadata_Well", how="inner")
        feats = [c for c in merged.columns if "Metadata" not in c]
        z = o_norm.mad_robustize(merged[feats].to_numpy(float), (merged["Metadata_Compound"] == "DMSO").to_numpy())
        assert list(got.columns) == [c for c in merged.columns if c not in feats] + ["Metadata_Timepoint"] + feats
        assert list(got["Metadata_Well"]) == list(merged["Metadata_Well"])
        np.testing.assert_allclose(got[feats].to_numpy(float), z, rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError, match="no GPU kernel"):
        nz.concatenate_csv_from_s3("b", ["P1"], ["24h"], "exp", "out", "DMSO", "norm", "sum", False, qc_drop, s3)


@pytest.mark.parametrize("agg", ["mean", "median"])
@pytest.mark.parametrize("qc_drop", [False, True])
def test_normalize_aggregation_matches_reference_golden(tmp_path, monkeypatch, golden_dir, agg, qc_drop):
    """tests/golden/well_agg.npz holds the per-well table the REFERENCE's concatenate_csv_from_s3 writes
    when pycytominer's normalize is the identity (oracle/make_golden.py): the drop-in's table plumbing and
    GPU aggregation (ips_well_mean_f64 / ips_well_median_f64) must reproduce it column by column, and hand
    the same control wells to the normalisation."""
    require_gpu()
    from image_processing_suite_b200.scripts import Normalize_CP_ami as nz, storage
    g = np.load(os.path.join(golden_dir, "well_agg.npz"))
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    for name in ("Image", "Nuclei", "Cells", "Cytoplasm"):
        s3.put_object(Bucket="b", Key=f"exp/P1/24h/{name}.csv", Body=g[f"in_{name}"].tobytes())
    s3.put_object(Bucket="b", Key="exp/Plate_P1_PlateMap.csv", Body=g["in_PlateMap"].tobytes())
    seen = {}

    def identity(profiles, features, control_mask):
        seen["controls"] = profiles.loc[np.asarray(control_mask, bool), "Metadata_Well"].tolist()
        meta = profiles[[c for c in profiles.columns if c not in features]].reset_index(drop=True)
        return pd.concat([meta, profiles[features].reset_index(drop=True)], axis=1)

    monkeypatch.setattr(nz, "normalize_mad_robustize", identity)
    keys = nz.concatenate_csv_from_s3("b", ["P1"], ["24h"], "exp", "out", "DMSO", "norm", agg, False, qc_drop, s3)
    got = pd.read_csv(io.BytesIO(s3.get_object(Bucket="out", Key=keys[0])["Body"].read()))
    key = f"{agg}_{'qc' if qc_drop else 'all'}"
    ref = pd.read_csv(io.BytesIO(g[f"out_{key}"].tobytes()))
    assert sorted(got.columns) == sorted(ref.columns)
    assert list(got["Metadata_Well"]) == list(ref["Metadata_Well"])
    for c in ref.columns:
        if c.startswith("Metadata"):
            assert list(got[c]) == list(ref[c]), c
        else:
            np.testing.assert_allclose(got[c].to_numpy(float), ref[c].to_numpy(float), rtol=1e-12, err_msg=c)
    assert seen["controls"] == g[f"controls_{key}"].tolist()


def test_feature_select_cosine_script(tmp_path, monkeypatch):
    """The cosine drop-in with an identity feature selection: double sigmoid and per-group mean
    cosine (all groups in one kernel call) against scikit-learn, group by group."""
    require_gpu()
    from sklearn.metrics.pairwise import cosine_similarity
    from image_processing_suite_b200.scripts import Feature_select_cosine_ami as fs, storage
    from oracle import normalize as o_norm
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    rng = np.random.default_rng(12)
    rows = []
    for plate in ("P1", "P2"):
        for tp in ("24h", "48h"):
            frame = []
            for comp, reps in (("DMSO", 6), ("CMP1", 4), ("CMP2", 1), ("CMP3", 3)):
                for r in range(reps):
                    d = {"Metadata_Compound": comp, "Metadata_ConcLevel": 1 + (r % 2 if comp == "CMP1" else 0),
                         "Metadata_Well": f"{comp}{r}", "Metadata_Plate": plate, "Metadata_Timepoint": tp}
                    d.update({f"DNA_f{j}": rng.normal(0, 3) for j in range(40)})
                    frame.append(d)
            df = pd.DataFrame(frame)
            df.loc[2, "DNA_f3"] = np.nan
            s3.put_object(Bucket="b", Key=f"exp/{plate}/Normalized_features_{tp}.csv", Body=df.to_csv(index=False).encode())
            rows.append(df)
    s3.put_object(Bucket="b", Key="exp/P1/sub/Normalized_features_x.csv", Body=b"ignored: deeper than the plate folder")
    ident = lambda profiles, features, **kw: profiles
    sims = fs.concatenate_normalized_csv_from_s3("b", ["P1", "P2"], "exp", False, "out", "res", "EXP", 0.5, 0.9,
                                                 feature_select=ident, s3=s3)
    allrows = pd.concat(rows, ignore_index=True)
    feats = [c for c in allrows.columns if "Metadata" not in c]
    allrows[feats] = np.abs(o_norm.double_sigmoid(allrows[feats].to_numpy(float)))
    dsig = pd.read_csv(io.BytesIO(s3.get_object(Bucket="out", Key="res/EXP_CP_features_selected_allTimes_dSig.csv")["Body"].read()))
    np.testing.assert_allclose(dsig[feats].to_numpy(float), allrows[feats].to_numpy(float), rtol=1e-12, equal_nan=True)
    expected = []
    for key in allrows[fs.GROUP_KEYS].drop_duplicates().values:
        grp = allrows[(allrows.Metadata_Compound == key[0]) & (allrows.Metadata_Timepoint == key[1]) &
                      (allrows.Metadata_ConcLevel == key[2])]
        s = cosine_similarity(grp[feats].fillna(0))
        v = s[np.triu_indices_from(s, k=1)]
        expected.append(np.mean(v) if len(v) else np.nan)
    assert list(sims[fs.GROUP_KEYS].itertuples(index=False, name=None)) == \
        [tuple(k) for k in allrows[fs.GROUP_KEYS].drop_duplicates().values]
    np.testing.assert_allclose(sims["average_cosine_similarity"].to_numpy(), np.asarray(expected), atol=1e-5, equal_nan=True)
    saved = pd.read_csv(io.BytesIO(s3.get_object(Bucket="out", Key="res/EXP_Average_cosine_similarity.csv")["Body"].read()))
    assert len(saved) == len(sims)


@pytest.mark.parametrize("per_time", [False, True])
def test_feature_select_cosine_script_matches_reference_golden(tmp_path, monkeypatch, golden_dir, per_time):
    """tests/golden/cosine_script.npz: the files the REFERENCE's concatenate_normalized_csv_from_s3 writes
    with an identity feature selection (oracle/make_golden.py).  The drop-in (ips_double_sigmoid_abs, all
    groups in one ips_cosine_triu call) writes the same tables: keys, order, values."""
    require_gpu()
    from image_processing_suite_b200.scripts import Feature_select_cosine_ami as fs, storage
    g = np.load(os.path.join(golden_dir, "cosine_script.npz"))
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    for plate in ("P1", "P2"):
        for tp in ("24h", "48h"):
            s3.put_object(Bucket="b", Key=f"exp/{plate}/Normalized_features_{tp}.csv", Body=g[f"in_{plate}_{tp}"].tobytes())
    s3.put_object(Bucket="b", Key="exp/P1/sub/Normalized_features_x.csv", Body=b"deeper than the plate folder: not listed")
    ident = lambda profiles, features, **kw: profiles
    fs.concatenate_normalized_csv_from_s3("b", ["P1", "P2"], "exp", per_time, "out", "res", "EXP", 0.5, 0.9,
                                          feature_select=ident, s3=s3)
    tag = "pertime" if per_time else "global"

    def both(name):
        got = pd.read_csv(io.BytesIO(s3.get_object(Bucket="out", Key=f"res/{name}")["Body"].read()))
        return got, pd.read_csv(io.BytesIO(g[f"{tag}_{name}"].tobytes()))

    for name, tol in (("EXP_CP_features_selected_allTimes_raw.csv", 0.0), ("EXP_CP_features_selected_allTimes_dSig.csv", 1e-12)):
        got, ref = both(name)
        assert list(got.columns) == list(ref.columns) and len(got) == len(ref)
        for c in ref.columns:
            if "Metadata" in c:
                assert list(got[c]) == list(ref[c]), c
            else:
                # pandas writes small floats with 15 decimals: the files hold the values to 1e-15 absolute
                np.testing.assert_allclose(got[c].to_numpy(float), ref[c].to_numpy(float), rtol=tol, atol=2e-15 if tol else 0.0,
                                           equal_nan=True, err_msg=c)
    got, ref = both("EXP_Average_cosine_similarity.csv")
    assert list(got.columns) == list(ref.columns)
    for c in fs.GROUP_KEYS:
        assert list(got[c]) == list(ref[c]), c
    np.testing.assert_allclose(got["average_cosine_similarity"].to_numpy(float), ref["average_cosine_similarity"].to_numpy(float),
                               atol=1e-5, equal_nan=True)


def test_pycyto_pertime_script(tmp_path, monkeypatch):
    """Pycyto_pertime drop-in with an identity feature selection against pandas / oracle /
    scikit-learn on the same tables (per-pair similarity vectors included)."""
    require_gpu()
    from functools import reduce
    from sklearn.metrics.pairwise import cosine_similarity
    from image_processing_suite_b200.scripts import Pycyto_pertime as pp, storage
    from oracle import normalize as o_norm
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    rng = np.random.default_rng(31)
    img_rows, tabs = [], {"Nuclei": [], "Cells": [], "Cytoplasm": []}
    n = 0
    for wi in range(12):
        comp = "DMSO" if wi % 3 == 0 else f"CMP{wi % 4}"
        for site in (1, 2):
            n += 1
            img_rows.append({"ImageNumber": n, "Metadata_Plate": "Plate_1", "Metadata_Site": site, "Metadata_Well": f"W{wi:02d}",
                             "Metadata_Timepoint": "24h", "Metadata_Compound": comp, "Metadata_ConcLevel": 1 + wi % 2,
                             "Count_Nuclei": int(rng.integers(10, 40)), "FileName_DNA": "x.tiff",
                             "Granularity_1": rng.normal(5, 1)})
            for name in tabs:
                for _ in range(int(rng.integers(3, 6))):
                    tabs[name].append({"ImageNumber": n, "AreaShape_Area": int(rng.integers(300, 900)),
                                       "Intensity_Mean_DNA": rng.normal(0.2, 0.05), "Texture_1": rng.normal(1, 0.3)})
    s3.put_object(Bucket="b", Key="proj/Plate_1/24h/Image.csv", Body=pd.DataFrame(img_rows).to_csv(index=False).encode())
    for name, rows in tabs.items():
        s3.put_object(Bucket="b", Key=f"proj/Plate_1/24h/{name}.csv", Body=pd.DataFrame(rows).to_csv(index=False).encode())
    ident = lambda profiles, features, **kw: profiles
    res = pp.concatenate_csv_from_s3("b", ["24h"], "proj/Plate_1", "out", "res", feature_select=ident, s3=s3)
    selected, averaged, sims = res["24h"]
    # ---- the same steps in pandas -------------------------------------------------------------
    image = pd.DataFrame(img_rows)
    objs = {k: pd.DataFrame(v).merge(image[["ImageNumber"] + pp.IMAGE_META], on="ImageNumber", how="left")
            .drop(["ImageNumber", "Metadata_Site", "Metadata_ConcLevel"], axis=1) for k, v in tabs.items()}
    image = image.drop(["ImageNumber"], axis=1)
    image = image.drop(columns=[c for c in image.columns
                                if not pd.api.types.is_numeric_dtype(image[c]) and not c.startswith("Metadata")])
    g = {k: v.groupby(pp.KEYS, as_index=False).mean() for k, v in objs.items()}
    image = image.groupby(pp.KEYS, as_index=False).mean().rename(columns=lambda x: "Image_" + x if x not in pp.IMAGE_META else x)
    merged = reduce(lambda l, r: pd.merge(l, r, on=pp.KEYS, how="outer"), [g["Cells"], g["Nuclei"], image, g["Cytoplasm"]])
    feats = [c for c in merged.columns if "Metadata" not in c]
    z = np.abs(o_norm.double_sigmoid(o_norm.mad_robustize(merged[feats].to_numpy(float), (merged.Metadata_Compound == "DMSO").to_numpy())))
    assert [c for c in selected.columns if "Metadata" not in c] == feats
    np.testing.assert_allclose(selected[feats].to_numpy(float), z, rtol=1e-8, atol=1e-10)
    merged[feats] = z
    k = 0
    for key in merged[["Metadata_Compound", "Metadata_Timepoint", "Metadata_ConcLevel"]].drop_duplicates().values:
        grp = merged[(merged.Metadata_Compound == key[0]) & (merged.Metadata_ConcLevel == key[2])]
        smat = cosine_similarity(grp[feats + ["Metadata_Site"]].fillna(0)) if False else cosine_similarity(
            grp.drop(columns=["Metadata_Plate", "Metadata_Well", "Metadata_Site", "Metadata_Compound", "Metadata_Timepoint",
                              "Metadata_ConcLevel"]).fillna(0))
        v = smat[np.triu_indices_from(smat, k=1)]
        assert averaged.loc[k, "Metadata_compound_code"] == key[0]
        np.testing.assert_allclose(sims.loc[k, "cosine_similarities"], v, atol=1e-5)
        if len(v):
            assert abs(averaged.loc[k, "average_cosine_similarity"] - v.mean()) < 1e-5
        else:
            assert np.isnan(averaged.loc[k, "average_cosine_similarity"])
        k += 1
    assert k == len(averaged)


def test_pycyto_pertime_script_matches_reference_golden(tmp_path, monkeypatch, golden_dir):
    """tests/golden/pycyto_pertime.npz: the three files the REFERENCE's Pycyto_pertime.concatenate_csv_from_s3
    writes when pycytominer's normalize and feature_select are the identity (oracle/make_golden.py).  The
    drop-in's four-key group means (ips_well_mean_f64), |double sigmoid| and per-pair cosine values
    (ips_cosine_triu_pairs) reproduce them, suffixes of the colliding column names included."""
    require_gpu()
    from image_processing_suite_b200.scripts import Pycyto_pertime as pp, storage
    g = np.load(os.path.join(golden_dir, "pycyto_pertime.npz"))
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    for name in ("Image", "Nuclei", "Cells", "Cytoplasm"):
        s3.put_object(Bucket="b", Key=f"proj/Plate_1/24h/{name}.csv", Body=g[f"in_{name}"].tobytes())

    def identity(profiles, features, control_mask):
        meta = profiles[[c for c in profiles.columns if c not in features]].reset_index(drop=True)
        return pd.concat([meta, profiles[features].reset_index(drop=True)], axis=1)

    monkeypatch.setattr(pp, "normalize_mad_robustize", identity)
    ident = lambda profiles, features, **kw: profiles
    pp.concatenate_csv_from_s3("b", ["24h"], "proj/Plate_1", "out", "res", feature_select=ident, s3=s3)

    def both(name):
        got = pd.read_csv(io.BytesIO(s3.get_object(Bucket="out", Key=f"res/24h/{name}")["Body"].read()))
        return got, pd.read_csv(io.BytesIO(g[name].tobytes()))

    got, ref = both("CP_features_selected.csv")
    assert sorted(got.columns) == sorted(ref.columns) and len(got) == len(ref)
    assert [c for c in got.columns if "Metadata" not in c] == [c for c in ref.columns if "Metadata" not in c]
    for c in ref.columns:
        if c in pp.KEYS:
            assert list(got[c]) == list(ref[c]), c
        else:
            np.testing.assert_allclose(got[c].to_numpy(float), ref[c].to_numpy(float), rtol=1e-12, atol=2e-15, err_msg=c)
    got, ref = both("CPfeatures_average_cosine_similarity.csv")
    assert list(got.columns) == list(ref.columns)
    for c in ("Metadata_compound_code", "Metadata_Timepoint", "Metadata_compound_concentration"):
        assert list(got[c]) == list(ref[c]), c
    np.testing.assert_allclose(got["average_cosine_similarity"].to_numpy(float), ref["average_cosine_similarity"].to_numpy(float),
                               atol=1e-5, equal_nan=True)
    got, ref = both("CPfeatures_cosine_similarities.csv")
    assert list(got.columns) == list(ref.columns) and len(got) == len(ref)
    for a, b in zip(got["cosine_similarities"], ref["cosine_similarities"]):
        va, vb = (np.array(x.strip("[]").split(), dtype=float) for x in (a, b))
        np.testing.assert_allclose(va, vb, atol=1e-5)
