"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN FUNCTIONS.

Run in the authoring container only (``python oracle/make_golden.py``): it imports
MaxProjection.py, Illumination_QC_mult.py and Image_re-binning.py from /root/reference
with ``boto3`` / ``imageio`` / ``tifffile`` replaced by in-memory stand-ins (those
packages are not installed and carry no arithmetic), feeds them seeded inputs and stores
inputs + outputs.  /root/reference does not exist on the GPU box, so the tests read only
the committed fixtures.  Nothing from the reference is copied into this repository;
the fixtures hold numbers, not code.
"""
import importlib.util
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ---- stand-ins for the I/O packages the reference imports at module top ------------
class _Store(dict):
    """Fake S3: keys -> bytes."""

    def get_object(self, Bucket, Key):
        return {"Body": io.BytesIO(self[(Bucket, Key)])}

    def upload_fileobj(self, fileobj, bucket, key):
        self[(bucket, key)] = fileobj.read()


def _npy_bytes(a):
    b = io.BytesIO()
    np.save(b, a)
    return b.getvalue()


def _install_stubs():
    boto3 = types.ModuleType("boto3")
    boto3.client = lambda *a, **k: None
    boto3.resource = lambda *a, **k: None
    imageio = types.ModuleType("imageio")
    imageio.imread = lambda f: np.load(io.BytesIO(f.read()))
    imageio.imwrite = lambda f, arr, format=None: np.save(f, arr)
    tifffile = types.ModuleType("tifffile")
    tifffile.imread = lambda path: np.load(path)
    tqdm = types.ModuleType("tqdm")
    tqdm.tqdm = lambda it, **k: it
    for name, mod in (("boto3", boto3), ("imageio", imageio), ("tifffile", tifffile)):
        sys.modules.setdefault(name, mod)
    try:
        import tqdm as _t  # noqa: F401
    except Exception:
        sys.modules["tqdm"] = tqdm


def _load(filename, modname):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, filename))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_plane(rng, h, w, lo=100, hi=4000, saturate=True):
    a = rng.integers(lo, hi, (h, w), dtype=np.uint16)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(6):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        r = rng.integers(4, 12)
        blob = 20000.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2.0 * r * r))
        a = np.minimum(a.astype(np.float64) + blob, 65535).astype(np.uint16)
    if saturate:
        a[rng.integers(0, h, 5), rng.integers(0, w, 5)] = 65535
    return a


def smooth_illum(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    r2 = ((yy - h / 2) / h) ** 2 + ((xx - w / 2) / w) ** 2
    return (1.0 + 0.5 * (1.0 - r2 / r2.max()) + 0.01 * rng.random((h, w))).astype(np.float64)


def main():
    os.makedirs(OUT, exist_ok=True)
    _install_stubs()
    mp = _load("MaxProjection.py", "ref_maxproj")
    qc = _load("Illumination_QC_mult.py", "ref_qc")
    rb = _load("Image_re-binning.py", "ref_rebin")
    rng = np.random.default_rng(20261018)

    # ---- A1: MaxProjection.max_projection through its real signature ---------------
    store = _Store()
    cases = {}
    for name, (z, h, w) in {"z3": (3, 48, 64), "z5": (5, 33, 47), "z1": (1, 16, 24)}.items():
        planes = [synth_plane(rng, h, w) for _ in range(z)]
        keys = [f"exp/Images/r01c01f01p{p:02d}-ch1.tiff" for p in range(z)]
        for k, p in zip(keys, planes):
            store[("bkt", k)] = _npy_bytes(p)
        mp.max_projection(keys, "bkt", store)
        out_key = mp.modify_imagepath(keys[0])
        assert "ImagesStacked" in out_key
        cases[f"{name}_in"] = np.stack(planes)
        cases[f"{name}_out"] = np.load(io.BytesIO(store[("bkt", out_key)]))
    try:
        bad = [f"exp/Images/bad{p}.tiff" for p in range(2)]
        store[("bkt", bad[0])] = _npy_bytes(np.zeros((4, 4), np.uint16))
        store[("bkt", bad[1])] = _npy_bytes(np.zeros((4, 5), np.uint16))
        mp.max_projection(bad, "bkt", store)
        cases["mismatch_raises"] = np.array(0)
    except ValueError:
        cases["mismatch_raises"] = np.array(1)
    cases["out_key"] = np.array(out_key)
    np.savez_compressed(os.path.join(OUT, "maxproj.npz"), **cases)

    # ---- A2 / A6 / A7: Illumination_QC_mult ------------------------------------------
    cases = {}
    tmp = "/tmp/ips_golden"
    os.makedirs(tmp, exist_ok=True)
    shapes = {"a": (96, 128), "b": (64, 64), "c": (40, 56), "small": (20, 30), "mid": (30, 36)}
    for name, (h, w) in shapes.items():
        img = synth_plane(rng, h, w)
        ill = smooth_illum(rng, h, w)
        path = os.path.join(tmp, f"{name}.npy")
        np.save(path, img)
        # process_site: (index, paths, channels, illum_cache) -> (index, dict)
        idx, res = qc.process_site((7, [path, path, os.path.join(tmp, "missing.npy")],
                                    ["chA", "chB", "chC"], [ill, None, None]))
        assert idx == 7 and res["QC_Error_chC"] == "File Not Found"
        corrected = img.astype(float) / ill
        labels, magsum, powsum = qc.rps(corrected)
        cases[f"{name}_img"] = img
        cases[f"{name}_illum"] = ill
        cases[f"{name}_rps_labels"] = np.asarray(labels)
        cases[f"{name}_rps_mag"] = np.asarray(magsum, dtype=np.float64)
        cases[f"{name}_rps_pow"] = np.asarray(powsum, dtype=np.float64)
        cases[f"{name}_slope_corr"] = np.array(res["ImageQuality_PowerLogLogSlope_chA"], np.float64)
        cases[f"{name}_pct_corr"] = np.array(res["ImageQuality_PercentMaximal_chA"], np.float64)
        cases[f"{name}_slope_raw"] = np.array(res["ImageQuality_PowerLogLogSlope_chB"], np.float64)
        cases[f"{name}_pct_raw"] = np.array(res["ImageQuality_PercentMaximal_chB"], np.float64)
    # shape-mismatched illum is silently ignored (:149-153)
    img = synth_plane(rng, 32, 40)
    path = os.path.join(tmp, "mm.npy")
    np.save(path, img)
    _, res = qc.process_site((0, [path], ["x"], [np.ones((8, 8))]))
    _, res_raw = qc.process_site((0, [path], ["x"], [None]))
    assert res == res_raw
    cases["mismatch_img"] = img
    cases["mismatch_pct"] = np.array(res["ImageQuality_PercentMaximal_x"])
    cases["mismatch_slope"] = np.array(res["ImageQuality_PowerLogLogSlope_x"])
    const = np.full((48, 48), 1234.0)
    m = qc.calculate_qc_metrics(const, "k")
    cases["const_slope"] = np.array(m["ImageQuality_PowerLogLogSlope_k"])
    cases["const_pct"] = np.array(m["ImageQuality_PercentMaximal_k"])
    np.savez_compressed(os.path.join(OUT, "illum_qc.npz"), **cases)

    # ---- A3: Image_re-binning.process_image_in_memory (TIFF bytes -> TIFF bytes) ------
    from PIL import Image
    cases = {}
    for name, (h, w, oh, ow) in {"x2": (128, 128, 64, 64), "x4": (120, 160, 30, 40),
                                 "odd": (90, 70, 41, 33), "full": (64, 96, 32, 48)}.items():
        if name == "full":
            img = rng.integers(0, 65536, (h, w), dtype=np.uint16)
            img[10:30, 20:50] = 65535          # plateau -> overshoot -> byte-clip quirk
            img[40:50, 60:80] = 0
        else:
            img = synth_plane(rng, h, w)
            img[5:15, 5:25] = 65535
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="tiff")
        out_bytes = rb.process_image_in_memory(buf.getvalue(), target_size=(ow, oh))
        out = np.asarray(Image.open(io.BytesIO(out_bytes)), dtype=np.uint16)
        assert out.shape == (oh, ow)
        cases[f"{name}_in"] = img
        cases[f"{name}_out"] = out
    np.savez_compressed(os.path.join(OUT, "rebin_lanczos.npz"), **cases)

    # ---- A8: sklearn cosine_similarity as called by the reference ---------------------
    from sklearn.metrics.pairwise import cosine_similarity
    cases = {}
    for name, (n, d) in {"g4": (4, 37), "g7": (7, 130), "g2": (2, 5), "g1": (1, 9)}.items():
        x = rng.normal(size=(n, d))
        if n > 2:
            x[1] = 0.0                                     # zero row stays zero
        s = cosine_similarity(x)
        iu = np.triu_indices_from(s, k=1)
        v = s[iu]
        cases[f"{name}_x"] = x
        cases[f"{name}_triu"] = v
        cases[f"{name}_mean"] = np.array(np.mean(v) if len(v) > 0 else np.nan)
    np.savez_compressed(os.path.join(OUT, "cosine.npz"), **cases)

    # ---- crops: the reference's own scale_to_8bit (Cellpose_GPU_s3fs.py:34-43) ------------
    cp = _load("Cellpose_GPU_s3fs.py", "ref_cellpose")
    cases = {}
    for name, shape in {"a": (40, 40), "b": (24, 36), "const": (16, 16), "tiny": (3, 5)}.items():
        img = rng.uniform(0.0, 9000.0, shape).astype(np.float32)
        img[rng.random(shape) < 0.6] = 0.0                 # masked-out pixels of a crop are zeros
        if name == "const":
            img[:] = 123.5
        cases[f"{name}_in"] = img
        cases[f"{name}_out"] = cp.scale_to_8bit(img)
    cases["box_size"] = np.array(cp.BOX_SIZE)
    np.savez_compressed(os.path.join(OUT, "crops.npz"), **cases)
    tiff_golden()
    well_agg_golden()
    cosine_script_golden()
    pycyto_golden()
    print("golden fixtures written to", os.path.normpath(OUT))
    for f in sorted(os.listdir(OUT)):
        print(" ", f, os.path.getsize(os.path.join(OUT, f)))


def tiff_golden():
    """File-level fixtures of Image_re-binning.process_image_in_memory: the TIFF bytes that go in
    and the LZW TIFF bytes the reference returns (own generator: the other fixtures keep their
    random streams)."""
    from PIL import Image
    _install_stubs()
    rb = _load("Image_re-binning.py", "ref_rebin_tiff")
    rng = np.random.default_rng(20261018)
    cases = {}
    specs = {"noise": (96, 128, 48, 64, None), "lzw_in": (80, 120, 40, 60, "tiff_lzw"),
             "flat": (200, 400, 100, 200, None), "tall": (700, 60, 350, 30, "tiff_lzw"),
             "ident": (90, 110, 90, 110, None)}
    for name, (h, w, oh, ow, comp) in specs.items():
        if name == "flat":
            img = np.full((h, w), 300, np.uint16)
            img[50:60, 100:300] = 65535
        else:
            img = np.clip(rng.normal(400, 40, (h, w)), 0, 65535).astype(np.uint16)
            img[5:12, 5:25] = 65535
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="tiff", **({"compression": comp} if comp else {}))
        out_bytes = rb.process_image_in_memory(buf.getvalue(), target_size=(ow, oh))
        cases[f"{name}_in"] = np.frombuffer(buf.getvalue(), np.uint8)
        cases[f"{name}_file"] = np.frombuffer(out_bytes, np.uint8)
        cases[f"{name}_size"] = np.array([ow, oh])
    cases["pillow_version"] = np.array(Image.__version__)
    np.savez_compressed(os.path.join(OUT, "tiff_lzw.npz"), **cases)


class _Body:
    def __init__(self, b):
        self.b = b

    def read(self):
        return self.b


class _MemoryS3:
    """The slice of the boto3 S3 client the tabular scripts use, over a dict {(bucket, key): bytes}."""

    def __init__(self, store):
        self.store = store

    def get_object(self, Bucket, Key):
        return {"Body": _Body(self.store[(Bucket, Key)])}

    def put_object(self, Bucket, Key, Body):
        self.store[(Bucket, Key)] = Body.encode() if isinstance(Body, str) else Body

    def list_objects_v2(self, Bucket, Prefix, Delimiter=None):
        keys = sorted(k for b, k in self.store if b == Bucket and k.startswith(Prefix)
                      and (Delimiter is None or Delimiter not in k[len(Prefix):]))
        return {"Contents": [{"Key": k} for k in keys]}


def _load_with_stubs(filename, modname, stubs):
    """Import a reference script with ``stubs`` ({module name: module}) in sys.modules for the duration."""
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        return _load(filename, modname)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def well_agg_golden():
    """The well-level aggregation of Normalize_CP_ami.concatenate_csv_from_s3 (:29-151) run by the
    REFERENCE's own code on small tables: boto3 is an in-memory bucket, pycytominer (absent) is stubbed
    with ``annotate`` = the plate-map merge and ``normalize`` = identity, so the CSV the reference
    writes is its annotated per-well table -- qc_drop row removal, integer rescaling, groupby(...).agg
    (mean | median), the outer merge of the four tables.  Stored: the input CSVs, the output CSV per
    configuration, and the ``samples`` query the reference hands to normalize with the wells it selects."""
    import pandas as pd
    bucket = {}
    seen = {}
    boto3 = types.ModuleType("boto3")
    boto3.client = lambda *a, **k: _MemoryS3(bucket)
    botocore = types.ModuleType("botocore")
    config = types.ModuleType("botocore.config")
    config.Config = lambda **k: None
    botocore.config = config
    pyc = types.ModuleType("pycytominer")

    def annotate(profiles, platemap, join_on):
        return platemap.merge(profiles, left_on=join_on[0], right_on=join_on[1], how="inner")

    def normalize(profiles, features, samples, method):
        seen["samples"], seen["method"] = samples, method
        seen["control_wells"] = profiles.query(samples)["Metadata_Well"].tolist()
        return profiles.copy()

    pyc.annotate, pyc.normalize = annotate, normalize
    ref = _load_with_stubs("Normalize_CP_ami.py", "ref_normalize",
                           {"boto3": boto3, "botocore": botocore, "botocore.config": config, "pycytominer": pyc})
    rng = np.random.default_rng(20261019)
    wells = [f"{r}{c:02d}" for r in "ABC" for c in range(1, 6)]
    rows_img, tabs = [], {"Nuclei": [], "Cells": [], "Cytoplasm": []}
    image_number = 0
    for w in wells:
        for site in range(1, 4 if w != "B03" else 3):          # one well has fewer sites
            image_number += 1
            rows_img.append({"ImageNumber": image_number, "Metadata_Well": w, "Metadata_Site": site, "Metadata_Plate": "P1",
                             "Count_Nuclei": int(rng.integers(5, 30)), "ImageQC_Blurry": int(image_number % 7 == 0),
                             "ExecutionTime_X": 1.5, "URL_DNA": "file:x", "Intensity_Mean": float(rng.normal(0.3, 0.05))})
            for name in tabs:
                for obj in range(int(rng.integers(3, 8))):
                    tabs[name].append({"ImageNumber": image_number, "ObjectNumber": obj + 1,
                                       "AreaShape_Area": int(rng.integers(300, 1500)),
                                       "Intensity_MeanIntensity_DNA": float(rng.normal(0.2, 0.03)),
                                       "Intensity_StdIntensity_DNA": float(rng.normal(0.02, 0.004)) if rng.random() > 0.1 else float("nan"),
                                       "Location_Center_X": float(rng.uniform(0, 2160))})
    inputs = {"Image": pd.DataFrame(rows_img).to_csv(index=False)}
    for name, rows in tabs.items():
        inputs[name] = pd.DataFrame(rows).to_csv(index=False)
    pm = pd.DataFrame({"Metadata_Well": wells, "Metadata_Plate": "P1", "Metadata_ConcLevel": 1,
                       "Metadata_Compound": ["dmso" if i % 4 == 0 else f"cmp{i}" for i in range(len(wells))]})
    inputs["PlateMap"] = pm.to_csv(index=False)
    for name in ("Image", "Nuclei", "Cells", "Cytoplasm"):
        bucket[("b", f"exp/P1/24h/{name}.csv")] = inputs[name].encode()
    bucket[("b", "exp/Plate_P1_PlateMap.csv")] = inputs["PlateMap"].encode()
    cases = {f"in_{k}": np.frombuffer(v.encode(), np.uint8) for k, v in inputs.items()}
    for agg in ("mean", "median"):
        for qc_drop in (False, True):
            ref.concatenate_csv_from_s3(bucket_name="b", plates=["P1"], times=["24h"], base_folder_path="exp",
                                        output_bucket="out", DMSO="DMSO", output_prefix="norm", well_agg_func=agg,
                                        no_time_subFolder=False, qc_drop=qc_drop)
            key = f"{agg}_{'qc' if qc_drop else 'all'}"
            cases[f"out_{key}"] = np.frombuffer(bucket[("out", "norm/P1/Normalized_features_24h.csv")], np.uint8)
            cases[f"controls_{key}"] = np.array(seen["control_wells"])
    cases["samples_query"] = np.array(seen["samples"])
    cases["method"] = np.array(seen["method"])
    cases["pandas_version"] = np.array(pd.__version__)
    np.savez_compressed(os.path.join(OUT, "well_agg.npz"), **cases)


def cosine_script_golden():
    """Feature_select_cosine_ami.concatenate_normalized_csv_from_s3 (:39-164) run by the REFERENCE's own
    code: boto3 is an in-memory bucket, pycytominer.feature_select (absent) is stubbed with the identity
    selection (it writes the profiles it was given), so the files the reference writes pin its double
    sigmoid (:26-27, :117-118), the fillna / grouping rules and the per-group mean cosine (:131-156)."""
    import pandas as pd
    bucket = {}
    boto3 = types.ModuleType("boto3")
    boto3.client = lambda *a, **k: _MemoryS3(bucket)
    pyc = types.ModuleType("pycytominer")

    def feature_select(profiles, features, samples, operation, output_file, output_type, na_cutoff=None, corr_threshold=None):
        profiles.to_csv(output_file, index=False)

    pyc.feature_select = feature_select
    ref = _load_with_stubs("Feature_select_cosine_ami.py", "ref_cosine_script", {"boto3": boto3, "pycytominer": pyc})
    rng = np.random.default_rng(20261020)
    cases = {}
    for plate in ("P1", "P2"):
        for tp in ("24h", "48h"):
            frame = []
            for comp, reps in (("DMSO", 5), ("CMP1", 4), ("CMP2", 1), ("CMP3", 3)):
                for r in range(reps):
                    d = {"Metadata_Compound": comp, "Metadata_ConcLevel": 1 + (r % 2 if comp == "CMP1" else 0),
                         "Metadata_Well": f"{comp}{r}", "Metadata_Plate": plate, "Metadata_Timepoint": tp}
                    # short cells: csv.Sniffer (read_csv_from_s3, :33-36) must see several rows in 1024 characters
                    d.update({f"f{j}": round(float(rng.normal(0, 3)), 3) for j in range(12)})
                    frame.append(d)
            df = pd.DataFrame(frame)
            df.loc[2, "f3"] = np.nan
            df.loc[7, "f0"] = 0.0
            text = df.to_csv(index=False)
            bucket[("b", f"exp/{plate}/Normalized_features_{tp}.csv")] = text.encode()
            cases[f"in_{plate}_{tp}"] = np.frombuffer(text.encode(), np.uint8)
    bucket[("b", "exp/P1/sub/Normalized_features_x.csv")] = b"deeper than the plate folder: not listed"
    import tempfile
    for per_time in (False, True):
        with tempfile.TemporaryDirectory() as tmp:
            ref.concatenate_normalized_csv_from_s3(bucket_name="b", plates=["P1", "P2"], base_folder_path="exp",
                                                   per_time=per_time, output_bucket="out", output_prefix="res", exp="EXP",
                                                   na_cutoff=0.5, corr_3hold=0.9, local_dir=tmp)
        tag = "pertime" if per_time else "global"
        for name in ("EXP_CP_features_selected_allTimes_raw.csv", "EXP_CP_features_selected_allTimes_dSig.csv",
                     "EXP_Average_cosine_similarity.csv"):
            cases[f"{tag}_{name}"] = np.frombuffer(bucket[("out", f"res/{name}")], np.uint8)
    np.savez_compressed(os.path.join(OUT, "cosine_script.npz"), **cases)


def pycyto_golden():
    """Pycyto_pertime.concatenate_csv_from_s3 (:29-172) run by the REFERENCE's own code: boto3 is an
    in-memory bucket, pycytominer's normalize and feature_select are the identity, so the three files it
    writes pin the four-key groupby means (:69-72), the merge order, |double sigmoid| (:92-93) and the
    per-group cosine values and means (:115-163).  (The Image table carries no free-text column: the
    reference's ``dtype == 'object'`` test (:65) does not catch ``str`` columns under current pandas.)"""
    import pandas as pd
    bucket = {}
    boto3 = types.ModuleType("boto3")
    boto3.client = lambda *a, **k: _MemoryS3(bucket)
    pyc = types.ModuleType("pycytominer")
    pyc.annotate = lambda *a, **k: None
    pyc.normalize = lambda profiles, features, samples, method: profiles.copy()

    def feature_select(profiles, features, samples, operation, output_file, output_type):
        profiles.to_csv(output_file, index=False)

    pyc.feature_select = feature_select
    ref = _load_with_stubs("Pycyto_pertime.py", "ref_pycyto", {"boto3": boto3, "pycytominer": pyc})
    rng = np.random.default_rng(20261021)
    img_rows, tabs = [], {"Nuclei": [], "Cells": [], "Cytoplasm": []}
    n = 0
    for wi in range(12):
        comp = "DMSO" if wi % 3 == 0 else f"CMP{wi % 4}"
        for site in (1, 2):
            n += 1
            img_rows.append({"ImageNumber": n, "Metadata_Plate": "Plate_1", "Metadata_Site": site, "Metadata_Well": f"W{wi:02d}",
                             "Metadata_Timepoint": "24h", "Metadata_Compound": comp, "Metadata_ConcLevel": 1 + wi % 2,
                             "Count_Nuclei": int(rng.integers(10, 40)), "Granularity_1": round(float(rng.normal(5, 1)), 4)})
            for name in tabs:
                for _ in range(int(rng.integers(3, 6))):
                    tabs[name].append({"ImageNumber": n, "AreaShape_Area": int(rng.integers(300, 900)),
                                       "Intensity_Mean_DNA": round(float(rng.normal(0.2, 0.05)), 5),
                                       "Texture_1": round(float(rng.normal(1, 0.3)), 4)})
    cases = {}
    for name, rows in [("Image", img_rows)] + list(tabs.items()):
        text = pd.DataFrame(rows).to_csv(index=False)
        bucket[("b", f"proj/Plate_1/24h/{name}.csv")] = text.encode()
        cases[f"in_{name}"] = np.frombuffer(text.encode(), np.uint8)
    import contextlib
    import tempfile
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        ref.concatenate_csv_from_s3(bucket_name="b", times=["24h"], base_folder_path="proj/Plate_1", output_bucket="out",
                                    output_prefix="res", local_dir=tmp)
    for name in ("CP_features_selected.csv", "CPfeatures_average_cosine_similarity.csv", "CPfeatures_cosine_similarities.csv"):
        cases[name] = np.frombuffer(bucket[("out", f"res/24h/{name}")], np.uint8)
    np.savez_compressed(os.path.join(OUT, "pycyto_pertime.npz"), **cases)


if __name__ == "__main__":
    if sys.argv[1:] == ["pycyto"]:
        os.makedirs(OUT, exist_ok=True)
        pycyto_golden()
    elif sys.argv[1:] == ["cosine_script"]:
        os.makedirs(OUT, exist_ok=True)
        cosine_script_golden()
    elif sys.argv[1:] == ["tiff"]:
        os.makedirs(OUT, exist_ok=True)
        tiff_golden()
    elif sys.argv[1:] == ["wellagg"]:
        os.makedirs(OUT, exist_ok=True)
        well_agg_golden()
    else:
        main()
