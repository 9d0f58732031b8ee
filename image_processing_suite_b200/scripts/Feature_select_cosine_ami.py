"""Drop-in for Feature_select_cosine_ami.py: same flags and output keys.

Concatenate the plates' ``Normalized_features*`` tables, run feature selection, apply the
double sigmoid, and report the mean strict-upper-triangle cosine similarity of every
(compound, timepoint, concentration) replicate group (Feature_select_cosine_ami.py:39-164).

On the GPU (libips.so): the double sigmoid (``ips_double_sigmoid_abs``, :26-27, :117-118) and
the cosine step -- ALL replicate groups in one ``ips_cosine_triu`` call instead of one
scikit-learn call per group (:131-156).  Feature selection (:56-109) is pycytominer's
``feature_select`` -- pandas logic outside the hot path: it is called
if pycytominer is installed, or the caller passes its own ``feature_select`` callable; it is
not re-implemented here and there is no silent substitute.

``bucket.list_objects_v2`` of the reference becomes ``storage`` listing of the prefix.
"""
import argparse
import logging
from io import StringIO

import numpy as np
import pandas as pd

from . import storage

logging.basicConfig(format='%(asctime)s - %(levelname)s - %(message)s', level=logging.INFO)
logger = logging.getLogger(__name__)

k = 3
alpha = 2.3538
GROUP_KEYS = ['Metadata_Compound', 'Metadata_Timepoint', 'Metadata_ConcLevel']


def double_sigmoid_abs(values):
    """|(x/alpha)^k / sqrt(1 + (x/alpha)^(2k))| of a float64 array, on the GPU."""
    import torch
    from .. import ops
    x = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).cuda()
    return ops.double_sigmoid_abs(x, k, alpha).cpu().numpy()


def read_csv_from_s3(bucket_name, file_key, s3=None):
    s3 = s3 or storage.client()
    content = s3.get_object(Bucket=bucket_name, Key=file_key)['Body'].read().decode('utf-8')
    return pd.read_csv(StringIO(content), sep=storage.sniff_delimiter(content))


def drop_rows_without_group(df):
    """Rows with a missing group key belong to no replicate group: the reference selects groups
    with ``==`` (Feature_select_cosine_ami.py:132-140), which never matches NaN."""
    missing = df[GROUP_KEYS].isna().any(axis=1)
    if missing.any():
        logger.warning("Dropping %d rows with a missing %s", int(missing.sum()), " / ".join(GROUP_KEYS))
        df = df[~missing]
    return df


def average_cosine_similarities(profiles):
    """One row per (compound, timepoint, concentration) group, in order of first appearance,
    with the mean of the strict upper triangle of the group's cosine-similarity matrix
    (NaN for a single-replicate group), Feature_select_cosine_ami.py:131-156."""
    import torch
    from .. import ops
    cos = profiles.drop(columns=['Metadata_Plate', 'Metadata_Well'])
    cos = drop_rows_without_group(cos)
    keys = cos[GROUP_KEYS].drop_duplicates()
    key_rows = list(keys.itertuples(index=False, name=None))
    # stable sort of the rows by group id of first appearance makes every group contiguous
    gid = pd.Series(range(len(key_rows)), index=pd.MultiIndex.from_tuples(key_rows))
    row_gid = gid.reindex(pd.MultiIndex.from_frame(cos[GROUP_KEYS])).to_numpy()
    assert not np.isnan(row_gid.astype(np.float64)).any(), "every row must belong to a group"
    order = np.argsort(row_gid, kind="stable")
    feats = cos.drop(columns=GROUP_KEYS).fillna(0).to_numpy(dtype=np.float64)[order]
    x = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32)).cuda()
    g = torch.from_numpy(row_gid[order].astype(np.int32)).cuda()
    s, npairs = ops.cosine_triu(x, g, n_groups=len(key_rows))
    s, npairs = s.cpu().numpy(), npairs.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        avg = np.where(npairs > 0, s / np.maximum(npairs, 1), np.nan)
    out = pd.DataFrame(key_rows, columns=GROUP_KEYS)
    out['average_cosine_similarity'] = avg
    return out


def _default_feature_select():
    try:
        from pycytominer import feature_select
    except ImportError as e:
        raise ImportError("feature selection is pycytominer.feature_select (not part of this package): install "
                          "pycytominer or pass feature_select=<callable> to concatenate_normalized_csv_from_s3") from e
    return feature_select


def _put_csv(s3, bucket, key, df):
    buf = StringIO()
    df.to_csv(buf, index=False)
    s3.put_object(Bucket=bucket, Key=key, Body=buf.getvalue().encode())
    logger.info(f"Saved s3://{bucket}/{key}")


def concatenate_normalized_csv_from_s3(bucket_name, plates, base_folder_path, per_time, output_bucket, output_prefix, exp,
                                       na_cutoff, corr_3hold, local_dir="temp_data", feature_select=None, s3=None):
    s3 = s3 or storage.client()
    res = storage.resource() if s3 is None or not hasattr(s3, "Bucket") else s3
    select = feature_select or _default_feature_select()
    ops_list = ["variance_threshold", "drop_na_columns", "correlation_threshold", "drop_outliers"]
    frames = []
    for plate in plates:
        prefix = f"{base_folder_path}/{plate}/"
        keys = [o.key for o in res.Bucket(bucket_name).objects.filter(Prefix=prefix)
                if 'Normalized_features' in o.key and '/' not in o.key[len(prefix):]]
        logger.info(f"Found {len(keys)} normalized feature files for plate {plate}")
        frames += [read_csv_from_s3(bucket_name, kk, s3) for kk in keys]
    normalized = pd.concat(frames, ignore_index=True)

    def run_select(df):
        features = df.columns[~df.columns.str.contains("Metadata")].tolist()
        return select(profiles=df, features=features, samples="all", na_cutoff=na_cutoff, corr_threshold=corr_3hold,
                      operation=ops_list)

    if per_time:
        parts = []
        for tp in normalized["Metadata_Timepoint"].unique():
            sel = run_select(normalized[normalized["Metadata_Timepoint"] == tp]).copy()
            sel["Metadata_Timepoint"] = tp
            parts.append(sel)
            _put_csv(s3, output_bucket, f"{output_prefix}/{exp}CP_features_selected_{tp}_dSig.csv", sel)
        selected = pd.concat(parts, ignore_index=True).fillna(0)
    else:
        selected = run_select(normalized)
    _put_csv(s3, output_bucket, f"{output_prefix}/{exp}_CP_features_selected_allTimes_raw.csv", selected)
    features = selected.columns[~selected.columns.str.contains("Metadata")].tolist()
    selected = selected.copy()
    selected[features] = double_sigmoid_abs(selected[features].to_numpy(dtype=np.float64))
    _put_csv(s3, output_bucket, f"{output_prefix}/{exp}_CP_features_selected_allTimes_dSig.csv", selected)
    sims = average_cosine_similarities(selected)
    _put_csv(s3, output_bucket, f"{output_prefix}/{exp}_Average_cosine_similarity.csv", sims)
    return sims


def build_parser():
    parser = argparse.ArgumentParser(description="Concatenate and normalize CellProfiler features from S3.")
    parser.add_argument("--bucket_name", type=str, required=True)
    parser.add_argument("--base_folder", type=str, required=True)
    parser.add_argument("--plates", nargs="+", required=True)
    parser.add_argument("--exp", type=str, required=True)
    parser.add_argument("--na_cutoff", type=float, default=0.5)
    parser.add_argument("--corr_3hold", type=float, default=0.9)
    parser.add_argument("--per_time", action='store_true')
    parser.add_argument("--output_bucket", type=str, required=True)
    parser.add_argument("--output_prefix", type=str, required=True)
    parser.add_argument("--local_dir", type=str, default="temp_data")
    return parser


if __name__ == "__main__":
    a = build_parser().parse_args()
    concatenate_normalized_csv_from_s3(bucket_name=a.bucket_name, base_folder_path=a.base_folder, plates=a.plates, exp=a.exp,
                                       na_cutoff=a.na_cutoff, corr_3hold=a.corr_3hold, per_time=a.per_time,
                                       output_bucket=a.output_bucket, output_prefix=a.output_prefix, local_dir=a.local_dir)
