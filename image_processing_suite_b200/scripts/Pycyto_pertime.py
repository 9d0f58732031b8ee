"""Drop-in for Pycyto_pertime.py (the older single-plate variant): same flags and output keys.

Per timepoint: read Image / Nuclei / Cells / Cytoplasm.csv, per-well means, robust-z
normalisation against the DMSO wells, |double sigmoid|, feature selection, then for every
(compound, timepoint, concentration) replicate group the mean cosine similarity AND the
vector of pairwise similarities (Pycyto_pertime.py:29-172).

GPU (libips.so): the groupby means (``ips_well_mean``, :69-72), ``mad_robustize`` (:84-89),
the double sigmoid (:92-93) and the cosine step with its per-pair vectors
(``ips_cosine_triu_pairs``, :120-156).  ``feature_select`` (:99-106) is pycytominer's and is
called as is / injected, exactly as in Feature_select_cosine_ami.py.
"""
import argparse
import logging
from functools import reduce
from io import StringIO

import numpy as np
import pandas as pd

from . import storage
from .Feature_select_cosine_ami import (GROUP_KEYS, _default_feature_select, double_sigmoid_abs,
                                         drop_rows_without_group)
from .Normalize_CP_ami import normalize_mad_robustize

logger = logging.getLogger(__name__)
KEYS = ['Metadata_Plate', 'Metadata_Well', 'Metadata_Timepoint', 'Metadata_Compound']
IMAGE_META = ['Metadata_Plate', 'Metadata_Site', 'Metadata_Well', 'Metadata_Timepoint', 'Metadata_Compound',
              'Metadata_ConcLevel']


def read_csv_from_s3(bucket_name, file_key, s3=None):
    s3 = s3 or storage.client()
    content = s3.get_object(Bucket=bucket_name, Key=file_key)['Body'].read().decode('utf-8')
    return pd.read_csv(StringIO(content), sep=storage.sniff_delimiter(content))


def group_mean_gpu(df, keys):
    """df.groupby(keys, as_index=False).mean() with the means computed by ips_well_mean."""
    import torch
    from .. import ops
    codes, uniques = pd.factorize(pd.MultiIndex.from_frame(df[keys]), sort=True)
    feats = [c for c in df.columns if c not in keys]
    vals = torch.from_numpy(np.ascontiguousarray(df[feats].to_numpy(dtype=np.float64))).cuda()
    ids = torch.from_numpy(codes.astype(np.int32)).cuda()
    means, _ = ops.well_mean_f64(vals, ids, len(uniques))
    out = pd.DataFrame(means.cpu().numpy(), columns=feats)
    key_df = uniques.to_frame(index=False)
    key_df.columns = keys
    return pd.concat([key_df, out], axis=1)


def group_cosine(selected):
    """(averages DataFrame, similarities DataFrame) of Pycyto_pertime.py:115-163."""
    import torch
    from .. import ops
    cos = selected.drop(columns=[c for c in ('Metadata_Plate', 'Metadata_Well', 'Metadata_Site') if c in selected.columns])
    cos = drop_rows_without_group(cos)
    key_rows = list(cos[GROUP_KEYS].drop_duplicates().itertuples(index=False, name=None))
    gid = pd.Series(range(len(key_rows)), index=pd.MultiIndex.from_tuples(key_rows))
    row_gid = gid.reindex(pd.MultiIndex.from_frame(cos[GROUP_KEYS])).to_numpy()
    assert not np.isnan(row_gid.astype(np.float64)).any(), "every row must belong to a group"
    order = np.argsort(row_gid, kind="stable")
    feats = cos.drop(columns=GROUP_KEYS).fillna(0).to_numpy(dtype=np.float64)[order]
    sizes = np.bincount(row_gid, minlength=len(key_rows)).tolist()
    s, npairs, pairs, offsets = ops.cosine_triu_pairs(
        torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32)).cuda(),
        torch.from_numpy(row_gid[order].astype(np.int32)).cuda(), group_sizes=sizes)
    s, npairs, pairs, offsets = s.cpu().numpy(), npairs.cpu().numpy(), pairs.cpu().numpy(), offsets.cpu().numpy()
    averaged, similarities = [], []
    index = cos.index.to_numpy()[order]
    start = 0
    for g, key in enumerate(key_rows):
        vals = pairs[offsets[g]:offsets[g + 1]]
        averaged.append({'Metadata_compound_code': key[0], 'Metadata_Timepoint': key[1],
                         'Metadata_compound_concentration': key[2],
                         'average_cosine_similarity': float(s[g] / npairs[g]) if npairs[g] > 0 else np.nan})
        similarities.append({'Metadata_Compound': key[0], 'Metadata_Timepoint': key[1], 'Metadata_Condition': key[2],
                             'Replicates': pd.Index(index[start:start + sizes[g]]), 'cosine_similarities': vals})
        start += sizes[g]
    return pd.DataFrame(averaged), pd.DataFrame(similarities)


def _put_csv(s3, bucket, key, df):
    buf = StringIO()
    df.to_csv(buf, index=False)
    s3.put_object(Bucket=bucket, Key=key, Body=buf.getvalue().encode())
    print(f"Saved to S3: s3://{bucket}/{key}")


def concatenate_csv_from_s3(bucket_name, times, base_folder_path, output_bucket, output_prefix, local_dir="temp_data",
                            feature_select=None, s3=None):
    s3 = s3 or storage.client()
    select = feature_select or _default_feature_select()
    results = {}
    for time in times:
        print(time)
        image = read_csv_from_s3(bucket_name, f"{base_folder_path}/{time}/Image.csv", s3)
        objs = {n: read_csv_from_s3(bucket_name, f"{base_folder_path}/{time}/{n}.csv", s3) for n in ("Nuclei", "Cells", "Cytoplasm")}
        if 'Metadata_Site' not in objs["Nuclei"].columns:
            for n in objs:
                objs[n] = objs[n].merge(image[['ImageNumber'] + IMAGE_META], on='ImageNumber', how='left')
        for n in objs:
            objs[n] = objs[n].drop(['ImageNumber', 'Metadata_Site', 'Metadata_ConcLevel'], axis=1)
        image = image.drop(['ImageNumber'], axis=1)
        # the reference tests `dtype == 'object'` (pandas 1.5: strings); newer pandas types strings
        # as `str`, so "not numeric" is the version-independent form of the same rule (:65)
        image = image.drop(columns=[c for c in image.columns
                                    if not pd.api.types.is_numeric_dtype(image[c]) and not c.startswith('Metadata')])
        nuclei, cells, cytoplasm = (group_mean_gpu(objs[n], KEYS) for n in ("Nuclei", "Cells", "Cytoplasm"))
        image = group_mean_gpu(image, KEYS)
        image = image.rename(columns=lambda x: 'Image_' + x if x not in IMAGE_META else x)
        merged = reduce(lambda l, r: pd.merge(l, r, on=KEYS, how='outer'), [cells, nuclei, image, cytoplasm])
        merged["Metadata_Timepoint"] = time
        merged.Metadata_Plate = base_folder_path.split("/")[-1]
        features = merged.columns[~merged.columns.str.contains("Metadata")].to_list()
        control = ((merged["Metadata_Compound"] == 'DMSO') & (merged["Metadata_Timepoint"] == time)).to_numpy()
        normalized = normalize_mad_robustize(merged, features, control)
        normalized[features] = double_sigmoid_abs(normalized[features].to_numpy(dtype=np.float64))
        feats = normalized.columns[~normalized.columns.str.contains("Metadata")].tolist()
        selected = select(profiles=normalized, features=feats, samples="all",
                          operation=["variance_threshold", "drop_na_columns", "correlation_threshold", "drop_outliers"])
        _put_csv(s3, output_bucket, f"{output_prefix}/{time}/CP_features_selected.csv", selected)
        averaged, similarities = group_cosine(selected)
        _put_csv(s3, output_bucket, f"{output_prefix}/{time}/CPfeatures_average_cosine_similarity.csv", averaged)
        _put_csv(s3, output_bucket, f"{output_prefix}/{time}/CPfeatures_cosine_similarities.csv", similarities)
        results[time] = (selected, averaged, similarities)
    return results


def build_parser():
    parser = argparse.ArgumentParser(description="Concatenate CSV files from S3 for multiple plates.")
    parser.add_argument("--bucket_name", required=True, help="S3 bucket containing the files.")
    parser.add_argument("--base_folder", required=True, help="Base folder path in S3 where experiment folders are stored.")
    parser.add_argument("--times", nargs="+", required=True, help="List of times list to process (prefix as they are from CP Feature extraction).")
    parser.add_argument("--output_bucket", required=True, help="S3 bucket where output files will be saved.")
    parser.add_argument("--output_prefix", required=True, help="Prefix for the output files in S3.")
    parser.add_argument("--local_dir", default="temp_data", help="Local directory for temporary storage.")
    return parser


if __name__ == "__main__":
    a = build_parser().parse_args()
    print(f"Processing Plate {a.base_folder}...")
    concatenate_csv_from_s3(bucket_name=a.bucket_name, base_folder_path=a.base_folder, times=a.times,
                            output_bucket=a.output_bucket, output_prefix=a.output_prefix, local_dir=a.local_dir)
