"""Drop-in for Normalize_CP_ami.py: same flags, same input tables, same output CSV
(``{output_prefix}/{plate}/Normalized_features_{time}.csv``).

Per plate and timepoint: read Image / Nuclei / Cells / Cytoplasm.csv, drop the rows of images
that failed QC (--qc_drop), prefix the feature columns, rescale integer-typed columns by
max_sites / sites_in_well (--qc_drop), aggregate per well, merge the four tables, attach the
plate map, robust-z normalise against the DMSO wells of the timepoint
(Normalize_CP_ami.py:29-151).  Table plumbing stays in pandas, as in the reference; the two
numeric steps run on the GPU through libips.so:

  * ``groupby("Metadata_Well").agg("mean")`` (:126)          -> ips_well_mean
  * ``pycytominer.normalize(method="mad_robustize")`` (:137-142) -> ips_mad_robustize

pycytominer is not a dependency here.  ``annotate`` is restated as the inner merge of the
plate map with the profiles on the well (plate-map columns first); ``normalize`` as RobustMAD
fitted on the rows selected by the ``samples`` query -- both PARITY UNPINNED (the reference
does not pin pycytominer and has no test for them).
--well_agg_func mean and median have GPU kernels; any other pandas aggregation name is refused
(there is no CPU fallback).
"""
import argparse
import logging
from functools import reduce
from io import StringIO

import numpy as np
import pandas as pd

from . import storage

logging.basicConfig(format='%(asctime)s - %(levelname)s - %(message)s', level=logging.INFO)
logger = logging.getLogger(__name__)

TABLE_PREFIX = {'Image': 'Image_', 'Nuclei': 'DNA_', 'Cells': 'Cell_', 'Cytoplasm': 'Cyto_'}
DROP_SUBSTRINGS = ['ExecutionTime', 'ModuleError', 'URL']


def read_csv_from_s3(bucket_name, file_key, s3):
    logger.info(f"Reading CSV from s3://{bucket_name}/{file_key}")
    content = s3.get_object(Bucket=bucket_name, Key=file_key)['Body'].read().decode('utf-8')
    return pd.read_csv(StringIO(content), sep=storage.sniff_delimiter(content))


AGG_FUNCS = ("mean", "median")


def well_agg_gpu(df, func="mean"):
    """df with a Metadata_Well column and numeric feature columns -> one row per well (sorted
    by well, as pandas groupby does), aggregated in float64 on the GPU: ``mean``
    (ips_well_mean_f64) or ``median`` (ips_well_median_f64); NaN skipped per column as pandas
    does, any number of columns."""
    import torch
    from .. import ops
    if func not in AGG_FUNCS:
        raise ValueError("--well_agg_func %r has no GPU kernel (have: %s); there is no CPU fallback"
                         % (func, ", ".join(AGG_FUNCS)))
    wells, inverse = np.unique(df["Metadata_Well"].to_numpy(), return_inverse=True)
    feats = [c for c in df.columns if c != "Metadata_Well"]
    vals = torch.from_numpy(np.ascontiguousarray(df[feats].to_numpy(dtype=np.float64))).cuda()
    ids = torch.from_numpy(inverse.astype(np.int32)).cuda()
    agg = ops.well_mean_f64 if func == "mean" else ops.well_median_f64
    res, _ = agg(vals, ids, len(wells))
    out = pd.DataFrame(res.cpu().numpy(), columns=feats)
    out.insert(0, "Metadata_Well", wells)
    return out


def well_mean_gpu(df):
    return well_agg_gpu(df, "mean")


def annotate(profiles, platemap, join_on="Metadata_Well"):
    """Plate-map metadata attached to the well profiles: inner merge, plate-map columns first."""
    return platemap.merge(profiles, on=join_on, how="inner")


def normalize_mad_robustize(profiles, features, control_mask):
    """(x - median_ctrl) / (1.4826 * MAD_ctrl + 1e-18) on the feature columns (GPU); metadata
    columns first, as pycytominer returns them."""
    import torch
    from .. import ops
    x = torch.from_numpy(np.ascontiguousarray(profiles[features].to_numpy(dtype=np.float64))).cuda()
    ctrl = torch.from_numpy(np.asarray(control_mask, dtype=np.uint8)).cuda()
    z = ops.mad_robustize(x, ctrl).cpu().numpy()
    meta = profiles[[c for c in profiles.columns if c not in features]].reset_index(drop=True)
    return pd.concat([meta, pd.DataFrame(z, columns=features)], axis=1)


def aggregate_table(df, prefix, qc_drop, well_agg_func):
    """One CellProfiler table -> per-well profile (Normalize_CP_ami.py:83-127)."""
    keep_meta = {'Metadata_Well', 'Metadata_Site'} if qc_drop else {'Metadata_Well'}
    df = df.drop(columns=[c for c in df.columns
                          if c == 'ImageNumber'
                          or (c.startswith('Metadata') and c not in keep_meta)
                          or any(sub in c for sub in DROP_SUBSTRINGS)])
    df = df.rename(columns=lambda x: prefix + x if not x.startswith('Metadata_') else x)
    if qc_drop:
        site_counts = df.groupby("Metadata_Well")["Metadata_Site"].nunique()
        scaling = (site_counts.max() / site_counts).rename("scaling_factor")
        df = df.merge(scaling, on="Metadata_Well")
        to_scale = [c for c in df.select_dtypes(include="integer").columns if not c.startswith("Metadata")]
        df[to_scale] = df[to_scale].multiply(df["scaling_factor"], axis=0)
        df = df.drop(columns=["scaling_factor", "Metadata_Site"])
    numeric = df.select_dtypes(include="number").columns.tolist()
    return well_agg_gpu(df[["Metadata_Well"] + [c for c in numeric if c != "Metadata_Well"]], well_agg_func)


def concatenate_csv_from_s3(bucket_name, plates, times, base_folder_path, output_bucket, DMSO, output_prefix,
                            well_agg_func, no_time_subFolder, qc_drop, s3=None):
    s3 = s3 or storage.client()
    written = []
    for plate in plates:
        logger.info(f"Processing plate ID: {plate}")
        platemap = read_csv_from_s3(bucket_name, f"{base_folder_path}/Plate_{plate.lstrip('binned/')}_PlateMap.csv", s3)
        platemap = platemap[['Metadata_Compound', 'Metadata_ConcLevel', 'Metadata_Well', 'Metadata_Plate']].copy()
        platemap["Metadata_Compound"] = platemap["Metadata_Compound"].apply(lambda x: str(x).upper())
        for time in times:
            logger.info(f"Processing timepoint: {time}")
            tables = {}
            for name in TABLE_PREFIX:
                key = f"{base_folder_path}/{plate}/{name}.csv" if no_time_subFolder else f"{base_folder_path}/{plate}/{time}/{name}.csv"
                tables[name] = read_csv_from_s3(bucket_name, key, s3)
            image_df = tables["Image"]
            failing = image_df.loc[image_df.filter(like='ImageQC_').any(axis=1), 'ImageNumber']
            for name, df in list(tables.items()):
                if 'Metadata_Well' not in df.columns:
                    df = df.merge(image_df[['ImageNumber', 'Metadata_Well', 'Metadata_Site']], on='ImageNumber', how='left')
                    tables[name] = df
                if qc_drop:
                    tables[name] = df[~df['ImageNumber'].isin(failing)]
            for name, prefix in TABLE_PREFIX.items():
                tables[name] = aggregate_table(tables[name], prefix, qc_drop, well_agg_func)
            merged = reduce(lambda l, r: pd.merge(l, r, on='Metadata_Well', how='outer'), tables.values())
            merged = annotate(merged, platemap)
            merged["Metadata_Timepoint"] = time
            features = merged.columns[~merged.columns.str.contains("Metadata")].to_list()
            control = ((merged["Metadata_Compound"] == DMSO) & (merged["Metadata_Timepoint"] == time)).to_numpy()
            normalized = normalize_mad_robustize(merged, features, control)
            normalized[features] = normalized[features].astype(float)
            buf = StringIO()
            normalized.to_csv(buf, index=False)
            out_key = f"{output_prefix}/{plate}/Normalized_features_{time}.csv"
            s3.put_object(Bucket=output_bucket, Key=out_key, Body=buf.getvalue().encode())
            logger.info(f"Saved to S3: s3://{output_bucket}/{out_key}")
            written.append(out_key)
    return written


def build_parser():
    parser = argparse.ArgumentParser(description="Normalize each timepoint of a project folder, outputs normalized tables against DMSO.")
    parser.add_argument("--bucket_name", type=str, required=True, help="S3 bucket containing the files.")
    parser.add_argument("--base_folder", type=str, required=True, help="Base folder path in S3 where experiment folders are stored.")
    parser.add_argument("--plates", nargs="+", required=True, help="List of plates list to process (prefix Plate/Time/csv).")
    parser.add_argument("--times", nargs="+", help="List of times to process (prefix Plate/Time/csv).")
    parser.add_argument("--DMSO", type=str, default="DMSO", help="DMSO nomenclature used to normalize in the plateMap.")
    parser.add_argument("--output_bucket", type=str, required=True, help="S3 bucket where output files will be saved.")
    parser.add_argument("--output_prefix", type=str, required=True, help="Prefix for the output files in S3.")
    parser.add_argument("--well_agg_func", type=str, default="mean", help="Function to aggregate at well level. Default mean.")
    parser.add_argument("--no_time_subFolder", action='store_true')
    parser.add_argument("--qc_drop", action='store_true')
    return parser


if __name__ == "__main__":
    a = build_parser().parse_args()
    logger.info(f"Starting normalization for base folder: {a.base_folder}")
    concatenate_csv_from_s3(bucket_name=a.bucket_name, base_folder_path=a.base_folder, plates=a.plates, times=a.times,
                            no_time_subFolder=a.no_time_subFolder, qc_drop=a.qc_drop, DMSO=a.DMSO,
                            output_bucket=a.output_bucket, output_prefix=a.output_prefix, well_agg_func=a.well_agg_func)
