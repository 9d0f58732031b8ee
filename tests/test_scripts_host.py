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
