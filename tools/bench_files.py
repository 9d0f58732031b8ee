#!/usr/bin/env python
"""File-level, plate-scale run of the two drop-in scripts that touch image files, next to the reference's
own functions on the same files:

  MaxProjection      C x Z LZW-TIFF planes per field in -> C projected TIFFs out (scripts.MaxProjection.run;
                     reference: imageio/Pillow decode + np.maximum.reduce + TIFF write, MaxProjection.py:33-52)
  Feature_extraction C projected TIFFs + a uint16 label TIFF per site in -> Nuclei.csv / Image.csv out
                     (scripts.Feature_extraction.run; reference: the CellProfiler measurements restated by the
                     oracle + pandas to_csv)

    python tools/bench_files.py [--sites 256] [--distinct 16] [--root /dev/shm/ips_files] [--cpu-sites 4]

Synthetic fields with cells on a dark, quiet background (camera offset 300 +- 3 counts), so the planes
compress like real ones.  ``--distinct`` different fields are rendered and hard-linked to ``--sites``
site entries (the files are read from the page cache / tmpfs: storage bandwidth is not what is measured).
Outputs are compared: projected TIFFs byte for byte, integer features exactly, float features to 1e-5.
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
C_, Z_, H_, W_, CELLS = 5, 3, 2160, 2160, 2000


def quiet_field(labels, seed):
    """raw [C][Z][H][W] uint16: cells (blobs with texture) on a dark background with little noise."""
    rng = np.random.default_rng(seed)
    n = int(labels.max())
    amp = np.r_[0.0, rng.uniform(500.0, 8000.0, n)]
    cell = amp[labels]
    raw = np.empty((C_, Z_, H_, W_), np.uint16)
    for c in range(C_):
        gain = rng.uniform(0.5, 2.0)
        for z in range(Z_):
            att = np.exp(-((z - 1.0) / 1.5) ** 2)
            tex = 1.0 + 0.1 * rng.standard_normal((H_, W_)).astype(np.float32)
            bg = 300.0 + 3.0 * rng.standard_normal((H_, W_)).astype(np.float32)
            raw[c, z] = np.clip(np.rint(bg + cell * gain * att * tex), 0, 65535).astype(np.uint16)
    return raw


def build_files(root, sites, distinct):
    """Writes the data set; returns (data-set CSV path for MaxProjection, LoadData CSV for Feature_extraction)."""
    import pandas as pd
    import torch
    from image_processing_suite_b200 import synth
    from image_processing_suite_b200.scripts import tiffio
    img = os.path.join(root, "bkt", "exp", "Images")
    os.makedirs(img, exist_ok=True)
    os.makedirs(os.path.join(root, "sets"), exist_ok=True)
    ratio = []
    for d in range(distinct):
        lab = synth.make_labels(H_, W_, CELLS, seed=900 + d)
        raw = quiet_field(lab, 900 + d)
        blobs = tiffio.encode_lzw_from_device(torch.from_numpy(raw.reshape(C_ * Z_, H_, W_)).cuda())   # the files a microscope PC writes
        for c in range(C_):
            for z in range(Z_):
                with open(os.path.join(img, f"d{d}_c{c}_z{z}.tiff"), "wb") as f:
                    f.write(blobs[c * Z_ + z])
        ratio.append(sum(len(b) for b in blobs) / raw.nbytes)
        tiffio.write(os.path.join(img, f"d{d}_mask.tiff"), lab.astype(np.uint16), "tiff_lzw")
    rows, load = [], []
    for s in range(sites):
        d = s % distinct
        for z in range(Z_):                               # plane-major, channel-minor (MaxProjection.py:86-90)
            for c in range(C_):
                name = f"s{s}_c{c}_z{z}.tiff"
                dst = os.path.join(img, name)
                if not os.path.exists(dst):
                    os.link(os.path.join(img, f"d{d}_c{c}_z{z}.tiff"), dst)
                rows.append({"ChannelName": f"ch{c}", "ChannelID": c, "Image_FileName": name, "Image_PathName": "exp/Images",
                             "FieldID": s, "PlaneID": z, "PlateID": "P1", "Row": 1, "Col": 1, "Timestamp": 0})
        mask = os.path.join(img, f"s{s}_mask.tiff")
        if not os.path.exists(mask):
            os.link(os.path.join(img, f"d{d}_mask.tiff"), mask)
        # Feature_extraction reads the PROJECTED images, which MaxProjection writes to .../ImagesStacked/ under the first plane's name
        stacked = os.path.join(root, "bkt", "exp", "ImagesStacked")
        site = {"Metadata_Well": "A%02d" % (s // 9 + 1), "Metadata_Site": s % 9 + 1}
        for c in range(C_):
            site[f"FileName_ch{c}"] = f"s{s}_c{c}_z0.tiff"
            site[f"PathName_ch{c}"] = stacked
        site["Objects_FileName_Nuclei"] = f"s{s}_mask.tiff"
        site["Objects_PathName_Nuclei"] = img
        load.append(site)
    pd.DataFrame(rows).to_csv(os.path.join(root, "sets", "plate.csv"), sep=";", index=False)
    pd.DataFrame(load).to_csv(os.path.join(root, "load.csv"), index=False)
    return float(np.mean(ratio))


def reference_site(root, s):
    """What the reference does for one field: decode the planes (Pillow / libtiff, as imageio does), z-max per
    channel, write the projections; then the per-object measurements (oracle) and a pandas CSV block."""
    import pandas as pd
    from image_processing_suite_b200.scripts import Feature_extraction as fe, tiffio
    from oracle import object_stats as o_obj
    img = os.path.join(root, "bkt", "exp", "Images")
    t0 = time.perf_counter()
    proj, files = [], []
    for c in range(C_):
        planes = [tiffio.read(os.path.join(img, f"s{s}_c{c}_z{z}.tiff")) for z in range(Z_)]
        mp = np.maximum.reduce(planes)                   # MaxProjection.py:45
        proj.append(mp)
        files.append(tiffio.encode(mp))                  # MaxProjection.py:47-48
    t1 = time.perf_counter()
    lab = tiffio.read(os.path.join(img, f"s{s}_mask.tiff"))
    ints, flts = o_obj.object_stats(lab, np.stack(proj), None, 1.0 / 65535.0)
    text = fe.rows_to_frame(s + 1, ints, flts, [f"ch{c}" for c in range(C_)]).to_csv(index=False)
    t2 = time.perf_counter()
    return files, (ints, flts), t1 - t0, t2 - t1, text


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--root", default="/dev/shm/ips_files")
    ap.add_argument("--cpu-sites", type=int, default=4)
    ap.add_argument("--batch", type=int, default=8, help="sites per Feature_extraction batch")
    ap.add_argument("--mp-batch", type=int, default=4, help="fields per MaxProjection batch")
    ap.add_argument("--threads", type=int, default=8, help="reader threads of Feature_extraction")
    ap.add_argument("--mp-threads", type=int, default=8, help="reader / writer threads of MaxProjection")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args(argv)
    import pandas as pd
    import torch
    from image_processing_suite_b200.scripts import Feature_extraction as fe, MaxProjection as mp_script, storage
    shutil.rmtree(a.root, ignore_errors=True)
    os.makedirs(a.root)
    os.environ["IPS_STORAGE_ROOT"] = a.root
    ratio = build_files(a.root, a.sites, a.distinct)
    torch.cuda.synchronize()
    in_bytes = sum(os.path.getsize(os.path.join(a.root, "bkt", "exp", "Images", f"s{s}_c{c}_z{z}.tiff"))
                   for s in range(a.sites) for c in range(C_) for z in range(Z_))
    s3 = storage.client()
    # ---- MaxProjection, the script's own loop ---------------------------------------------------------
    mp_script.run("sets", "plate.csv", C_, Z_, "bkt", s3, batch_fields=a.mp_batch, threads=a.mp_threads)   # warm-up pass (page-locked slots, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_written = mp_script.run("sets", "plate.csv", C_, Z_, "bkt", s3, batch_fields=a.mp_batch, threads=a.mp_threads)
    torch.cuda.synchronize()
    t_mp = time.perf_counter() - t0
    assert n_written == a.sites * C_, n_written
    # ---- Feature_extraction over the projections ------------------------------------------------------
    out = os.path.join(a.root, "cp_out")
    fe.run(os.path.join(a.root, "load.csv"), out, batch=a.batch, threads=a.threads)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fe.run(os.path.join(a.root, "load.csv"), out, batch=a.batch, threads=a.threads)
    torch.cuda.synchronize()
    t_fe = time.perf_counter() - t0
    # ---- the reference's functions on the same files, one core, a few sites ---------------------------------
    nuclei = pd.read_csv(os.path.join(out, "Nuclei.csv"))
    t_ref_mp = t_ref_fe = 0.0
    for s in range(a.cpu_sites):
        files, (ints, flts), dt_mp, dt_fe, _ = reference_site(a.root, s)
        t_ref_mp += dt_mp
        t_ref_fe += dt_fe
        for c in range(C_):
            got = open(os.path.join(a.root, "bkt", "exp", "ImagesStacked", f"s{s}_c{c}_z0.tiff"), "rb").read()
            assert got == files[c], "projected TIFF of site %d channel %d differs" % (s, c)
        sub = nuclei[nuclei.ImageNumber == s + 1]
        assert len(sub) == ints.shape[0]
        np.testing.assert_array_equal(sub.ObjectNumber, ints[:, 0])
        np.testing.assert_array_equal(sub.AreaShape_Area, ints[:, 1])
        np.testing.assert_array_equal(sub.AreaShape_BoundingBoxMinimum_X, ints[:, 3])
        np.testing.assert_array_equal(sub.AreaShape_BoundingBoxMaximum_Y, ints[:, 4])
        np.testing.assert_allclose(sub.AreaShape_Center_X, flts[:, 1], rtol=1e-5)
        for c in range(C_):
            np.testing.assert_allclose(sub[f"Intensity_IntegratedIntensity_ch{c}"], flts[:, 2 + 5 * c], rtol=1e-5)
            np.testing.assert_allclose(sub[f"Intensity_MeanIntensity_ch{c}"], flts[:, 3 + 5 * c], rtol=1e-5)
            np.testing.assert_allclose(sub[f"Intensity_StdIntensity_ch{c}"], flts[:, 4 + 5 * c], rtol=1e-5,
                                       atol=1e-5 * float(np.abs(flts[:, 3 + 5 * c]).max()))
            np.testing.assert_allclose(sub[f"Intensity_MaxIntensity_ch{c}"], flts[:, 6 + 5 * c], rtol=1e-5)
    line = {"what": "file-level drop-in scripts at plate scale (TIFF files in -> TIFF / CSV files out), one GPU",
            "sites": a.sites, "distinct_fields": a.distinct, "lzw_bytes_per_pixel_byte": ratio,
            "maxprojection_fields_per_s": a.sites / t_mp, "maxprojection_input_gbs": in_bytes / t_mp / 1e9,
            "feature_extraction_fields_per_s": a.sites / t_fe,
            "reference_maxprojection_fields_per_s_1core": a.cpu_sites / t_ref_mp,
            "reference_features_fields_per_s_1core": a.cpu_sites / t_ref_fe,
            "host_cores": os.cpu_count(), "batch": a.batch, "threads": a.threads, "mp_batch": a.mp_batch, "mp_threads": a.mp_threads,
            "bytes_per_field": {"maxprojection_files_in": in_bytes // a.sites, "maxprojection_files_out": C_ * H_ * W_ * 2,
                                "feature_extraction_files_in": (C_ + 1) * H_ * W_ * 2},
            "limit": "MaxProjection moves 157 MB of file bytes per field through the host (read LZW planes, write "
                     "uncompressed projections as the reference does): bound by the host side (file copies on 16 cores and "
                     "the launching thread's share of the interpreter), not by the device",
            "outputs": "projected TIFFs byte-identical, integer features exact, float features within 1e-5 (%d sites compared)" % a.cpu_sites}
    if argv is None:
        print(json.dumps(line))
    if not a.keep:
        shutil.rmtree(a.root, ignore_errors=True)
    return line


if __name__ == "__main__":
    main()
