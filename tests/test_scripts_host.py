"""Host-side logic of the drop-in scripts (no GPU): flags, path rules, storage stand-in,
TIFF round trip, CSV sniffing -- mirrored from the reference scripts."""
import io

import numpy as np
import pytest

from image_processing_suite_b200.scripts import (Feature_extraction, Illumination_QC_mult, Image_rebinning,
                                                  MaxProjection, storage, tiffio)


def _flags(parser):
    return {a.option_strings[0]: (a.required, a.default) for a in parser._actions if a.option_strings and a.dest != 'help'}


def test_cli_flags_match_the_reference():
    # MaxProjection.py:56-61 -- five required flags
    f = _flags(MaxProjection.build_parser())
    assert set(f) == {"--bucket_data_set", "--data_set", "--channels", "--planes", "--bucket_images"}
    assert all(req for req, _ in f.values())
    # Image_re-binning.py:69-71
    f = _flags(Image_rebinning.build_parser())
    assert f == {"--bucket_name": (True, None), "--image_folder": (True, None), "--resolution": (False, 1080)}
    # Illumination_QC_mult.py:19-24
    a = Illumination_QC_mult.parse_args(["--load-data", "x.csv", "--data-path", "d", "--channels", "A", "B"])
    assert (a.load_data, a.data_path, a.illum_path, a.channels, a.output, a.threads) == \
        ("x.csv", "d", None, ["A", "B"], "QC_Results.csv", 24)
    # the CellProfiler command line of Feature_extraction_opt.py:166-167
    a = Feature_extraction.parse_args(["-c", "-r", "-p", "pipe.cppipe", "-o", "/out", "--data-file", "ld.csv"])
    assert (a.c, a.r, a.pipeline, a.output, a.data_file) == (True, True, "pipe.cppipe", "/out", "ld.csv")


def test_modify_imagepath_rules():
    # MaxProjection.py:16-22: only an exact 'Images' path component is replaced
    assert MaxProjection.modify_imagepath("p/Images/r01c01.tiff") == "p/ImagesStacked/r01c01.tiff"
    assert MaxProjection.modify_imagepath("p/MyImages/x.tiff") == "p/MyImages/x.tiff"
    assert MaxProjection.modify_imagepath("Images/a/Images/x") == "ImagesStacked/a/Images/x"


def test_local_storage_and_csv_sniffing(tmp_path, monkeypatch):
    monkeypatch.setenv("IPS_STORAGE_ROOT", str(tmp_path))
    s3 = storage.client()
    s3.upload_fileobj(io.BytesIO(b"PlateID;Image_FileName;Image_PathName\nP1;a.tif;x/Images\n"), "bkt", "sets/d.csv")
    df = MaxProjection.read_csv_from_s3("bkt", "sets/d.csv", s3)
    assert list(df.columns) == ["PlateID", "Image_FileName", "Image_PathName"] and df.iloc[0].Image_FileName == "a.tif"
    s3.put_object(Bucket="bkt", Key="e/Image/w1.tiff", Body=b"1")
    s3.put_object(Bucket="bkt", Key="e/Image_binned/w1.tiff", Body=b"2")
    keys = [o.key for o in storage.resource().Bucket("bkt").objects.filter(Prefix="e/Image/")]
    assert keys == ["e/Image/w1.tiff"]                      # trailing '/' keeps Image_binned out (:37-39)
    with pytest.raises(FileNotFoundError):
        s3.get_object(Bucket="bkt", Key="nope")


def test_tiff_round_trip_16bit():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 65536, (33, 47), dtype=np.uint16)
    for comp in (None, "tiff_lzw"):
        b = tiffio.decode(tiffio.encode(a, comp))
        assert b.dtype == np.uint16
        np.testing.assert_array_equal(a, b)


def test_feature_extraction_schema_helpers():
    import pandas as pd
    df = pd.DataFrame(columns=["FileName_DNA", "PathName_DNA", "FileName_IllumDNA", "FileName_ER",
                               "Objects_FileName_Nuclei", "Metadata_Well"])
    assert Feature_extraction.discover(df) == (["DNA", "ER"], ["DNA"], ["Nuclei"])
    ints = np.array([[3, 12, 2, 5, 6, 9]])
    flts = np.array([[3.5, 6.25, 10, 1, 0.5, 0.1, 2.0]])
    fr = Feature_extraction.rows_to_frame(7, ints, flts, ["DNA"])
    assert list(fr.columns) == Feature_extraction.object_columns(["DNA"])
    r = fr.iloc[0]
    assert (r.ImageNumber, r.ObjectNumber, r.AreaShape_Area) == (7, 3, 12)
    assert (r.AreaShape_BoundingBoxMinimum_X, r.AreaShape_BoundingBoxMaximum_Y) == (5, 6)
    assert (r.AreaShape_Center_X, r.Location_Center_Y) == (6.25, 3.5)
    text = fr.to_csv(index=False).splitlines()[1]
    assert text.startswith("7,3,12,5,2,9,6,")               # integer columns without a decimal point


def test_normalize_cli_flags_match_the_reference():
    from image_processing_suite_b200.scripts import Normalize_CP_ami
    f = _flags(Normalize_CP_ami.build_parser())          # Normalize_CP_ami.py:156-165
    assert set(f) == {"--bucket_name", "--base_folder", "--plates", "--times", "--DMSO", "--output_bucket",
                      "--output_prefix", "--well_agg_func", "--no_time_subFolder", "--qc_drop"}
    assert f["--DMSO"] == (False, "DMSO") and f["--well_agg_func"] == (False, "mean")
    assert f["--qc_drop"] == (False, False) and f["--plates"][0] and not f["--times"][0]


def test_feature_select_cosine_cli_flags_match_the_reference():
    from image_processing_suite_b200.scripts import Feature_select_cosine_ami as fs
    f = _flags(fs.build_parser())                          # Feature_select_cosine_ami.py:169-178
    assert f == {"--bucket_name": (True, None), "--base_folder": (True, None), "--plates": (True, None),
                 "--exp": (True, None), "--na_cutoff": (False, 0.5), "--corr_3hold": (False, 0.9),
                 "--per_time": (False, False), "--output_bucket": (True, None), "--output_prefix": (True, None),
                 "--local_dir": (False, "temp_data")}
    assert (fs.k, fs.alpha) == (3, 2.3538)
    with pytest.raises(ImportError, match="pycytominer"):
        fs._default_feature_select()


def test_pycyto_pertime_cli_flags_match_the_reference():
    from image_processing_suite_b200.scripts import Pycyto_pertime
    f = _flags(Pycyto_pertime.build_parser())              # Pycyto_pertime.py:179-184
    assert f == {"--bucket_name": (True, None), "--base_folder": (True, None), "--times": (True, None),
                 "--output_bucket": (True, None), "--output_prefix": (True, None), "--local_dir": (False, "temp_data")}


# ---- same-named entry points: `python <reference name>.py <reference flags>` ------------
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_NAMES = ["MaxProjection.py", "Illumination_QC_mult.py", "Image_re-binning.py", "Feature_extraction_opt.py",
                   "Normalize_CP_ami.py", "Feature_select_cosine_ami.py", "Pycyto_pertime.py"]


def _run(name, *argv, env=None, cwd=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "scripts", name), *argv], capture_output=True, text=True,
                          env=e, cwd=cwd, timeout=300)


def test_every_reference_script_name_exists_and_answers_help():
    for name in REFERENCE_NAMES:
        assert os.path.exists(os.path.join(ROOT, "scripts", name)), name
    flags = {"MaxProjection.py": ["--bucket_data_set", "--data_set", "--channels", "--planes", "--bucket_images"],
             "Illumination_QC_mult.py": ["--load-data", "--data-path", "--illum-path", "--channels", "--output", "--threads"],
             "Image_re-binning.py": ["--bucket_name", "--image_folder", "--resolution"],
             "Normalize_CP_ami.py": ["--well_agg_func", "--no_time_subFolder", "--qc_drop", "--DMSO"],
             "Feature_select_cosine_ami.py": ["--na_cutoff", "--corr_3hold", "--per_time", "--exp"],
             "Pycyto_pertime.py": ["--times", "--local_dir"]}
    for name, want in flags.items():
        r = _run(name, "--help")
        assert r.returncode == 0, r.stderr
        for f in want:
            assert f in r.stdout, (name, f)


def test_reference_command_lines_run_verbatim_against_local_storage(tmp_path):
    """The reference's own invocations (README / argparse defaults) against the directory stand-in
    for S3; inputs chosen so that no image reaches the GPU (this test runs without one)."""
    env = {"IPS_STORAGE_ROOT": str(tmp_path)}
    os.makedirs(tmp_path / "iric" / "exp" / "Image")
    (tmp_path / "iric" / "exp" / "Image" / "notes.txt").write_text("not an image")
    r = _run("Image_re-binning.py", "--bucket_name", "iric", "--image_folder", "exp/Image", "--resolution", "540", env=env)
    assert r.returncode == 0 and "Processed 0 images" in (r.stdout + r.stderr)
    os.makedirs(tmp_path / "sets")
    (tmp_path / "sets" / "d.csv").write_text("ChannelName;ChannelID;Image_FileName;Image_PathName;FieldID;PlaneID;PlateID;Row;Col;Timestamp\n"
                                             "DNA;1;a.tiff;p/Images;1;1;P1;1;1;0\n")
    r = _run("MaxProjection.py", "--bucket_data_set", "sets", "--data_set", "d.csv", "--channels", "5", "--planes", "3",
             "--bucket_images", "iric", env=env)
    assert r.returncode == 0 and "Skipping incomplete chunk" in (r.stdout + r.stderr)
    # Feature_extraction_opt.py has no flags: module constants + run_batch_processing() (:45-67, :183-184)
    r = _run("Feature_extraction_opt.py", env={**env, "IPS_PLATES_TO_RUN": "P01", "IPS_TIMES_TO_RUN": "6"})
    assert r.returncode == 0 and "missing" in (r.stdout + r.stderr) and "load_data_P01_6_illum.csv" in (r.stdout + r.stderr)


def test_feature_extraction_opt_keeps_the_reference_constants():
    import importlib.util
    spec = importlib.util.spec_from_file_location("feo", os.path.join(ROOT, "scripts", "Feature_extraction_opt.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)                                  # no S3 listing at import, unlike the reference (:65)
    for const in ("FOLDER", "CCPIPE_NAME", "PLATES_TO_RUN", "TIMES_TO_RUN", "BATCH_SIZE"):
        assert hasattr(m, const)
    assert m.TIMES_TO_RUN == ['6', '12', '24', '48'] and callable(m.run_batch_processing)
    argv = m.job_command("P02", "12")
    assert argv[:3] == ["-c", "-r", "-p"] and argv[-2] == "--data-file" and argv[-1].endswith("load_data_P02_12_illum.csv")


def test_qc_shell_is_written_here_and_can_bind_the_reference(tmp_path):
    """load_illum_cache finds either file name; bind_reference substitutes the arithmetic in a copy
    of a reference-shaped module without touching its shell."""
    np.save(tmp_path / "A_illum.npy", np.ones((2, 2)))
    np.save(tmp_path / "IllumB.npy", np.full((2, 2), 2.0))
    cache = Illumination_QC_mult.load_illum_cache(str(tmp_path), ["A", "B", "C"])
    assert cache[0][0, 0] == 1.0 and cache[1][0, 0] == 2.0 and cache[2] is None
    assert Illumination_QC_mult.load_illum_cache(None, ["A"]) == [None]
    stub = tmp_path / "ref.py"
    stub.write_text("def rps(img):\n    return 'cpu'\n\ndef process_site(t):\n    return 'cpu'\n\n"
                    "def calculate_saturation_cp_exact(i, mask=None):\n    return 'cpu'\n\n"
                    "def calculate_qc_metrics(i, c):\n    return 'cpu'\n\ndef main():\n    return process_site(None)\n")
    bound = Illumination_QC_mult.bind_reference(str(stub))
    assert bound.rps is Illumination_QC_mult.rps and bound.process_site is Illumination_QC_mult.process_site
