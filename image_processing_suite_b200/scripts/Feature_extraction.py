"""Drop-in for the CellProfiler command line that Feature_extraction_opt.py:166-167 runs:

    cellprofiler -c -r -p <pipeline.cppipe> -o <outdir> --data-file <load_data.csv>

Same flags, same outputs (``<outdir>/Image.csv`` and one ``<Object>.csv`` per label set, keyed
by ImageNumber / ObjectNumber, CellProfiler column names -- the schema Normalize_CP_ami.py:47-127
and Pycyto_pertime.py:46-75 consume).  The per-object arithmetic (MeasureObjectSizeShape /
MeasureObjectIntensity: area, bounding box, centre, integrated / mean / std / min / max
intensity per channel) runs in ips_field_fused on the GPU over the label masks named by the
``Objects_FileName_<Object>`` columns of the LoadData CSV (Cellpose masks, 0 = background).
The pipeline file is accepted and ignored: the measurement set is fixed.

LoadData columns used: FileName_<ch> / PathName_<ch> (images), FileName_Illum<ch> /
PathName_Illum<ch> (optional .npy illumination functions, plate-constant),
Objects_FileName_<Object> / Objects_PathName_<Object> (label images), Metadata_*.
Integer feature columns are written without a decimal point (Normalize_CP_ami.py:106-112
treats integer-typed columns specially).
"""
import argparse
import logging
import os

import numpy as np
import pandas as pd

from . import tiffio

logger = logging.getLogger(__name__)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="CellProfiler-compatible per-object feature extraction on the GPU")
    p.add_argument('-c', action='store_true', help="headless (accepted for compatibility)")
    p.add_argument('-r', action='store_true', help="run (accepted for compatibility)")
    p.add_argument('-p', '--pipeline', type=str, default=None, help="pipeline file (accepted, not interpreted)")
    p.add_argument('-o', '--output-directory', dest='output', type=str, required=True)
    p.add_argument('--data-file', type=str, required=True, help="LoadData CSV")
    p.add_argument('--image-directory', '-i', dest='image_dir', type=str, default=None,
                   help="base folder when the CSV has no PathName_ columns")
    p.add_argument('--batch', type=int, default=8, help="sites per device batch")
    p.add_argument('--raw-intensities', action='store_true',
                   help="report raw counts instead of CellProfiler's [0, 1] scaling (x / 65535)")
    return p.parse_args(argv)


def discover(df):
    """(image channels, illum channels, object names) from the LoadData column names."""
    chans = [c[len('FileName_'):] for c in df.columns if c.startswith('FileName_') and not c.startswith('FileName_Illum')]
    illum = [c for c in chans if f'FileName_Illum{c}' in df.columns]
    objs = [c[len('Objects_FileName_'):] for c in df.columns if c.startswith('Objects_FileName_')]
    return chans, illum, objs


def _path(row, prefix, name, base):
    fn = row[f'{prefix}FileName_{name}']
    pn_col = f'{prefix}PathName_{name}'
    folder = row[pn_col] if pn_col in row.index and isinstance(row[pn_col], str) else (base or '')
    return os.path.join(folder, fn)


def object_columns(channels):
    cols = ['ImageNumber', 'ObjectNumber', 'AreaShape_Area', 'AreaShape_BoundingBoxMinimum_X',
            'AreaShape_BoundingBoxMinimum_Y', 'AreaShape_BoundingBoxMaximum_X', 'AreaShape_BoundingBoxMaximum_Y',
            'AreaShape_Center_X', 'AreaShape_Center_Y', 'Location_Center_X', 'Location_Center_Y']
    for ch in channels:
        cols += [f'Intensity_IntegratedIntensity_{ch}', f'Intensity_MeanIntensity_{ch}',
                 f'Intensity_StdIntensity_{ch}', f'Intensity_MinIntensity_{ch}', f'Intensity_MaxIntensity_{ch}']
    return cols


def rows_to_frame(image_number, ints, flts, channels):
    """Device rows of one site -> CellProfiler-named DataFrame (ObjectNumber = label)."""
    n = ints.shape[0]
    d = {
        'ImageNumber': np.full(n, image_number, np.int64),
        'ObjectNumber': ints[:, 0].astype(np.int64),
        'AreaShape_Area': ints[:, 1].astype(np.int64),
        'AreaShape_BoundingBoxMinimum_X': ints[:, 3].astype(np.int64),
        'AreaShape_BoundingBoxMinimum_Y': ints[:, 2].astype(np.int64),
        'AreaShape_BoundingBoxMaximum_X': ints[:, 5].astype(np.int64),
        'AreaShape_BoundingBoxMaximum_Y': ints[:, 4].astype(np.int64),
        'AreaShape_Center_X': flts[:, 1].astype(np.float64),
        'AreaShape_Center_Y': flts[:, 0].astype(np.float64),
        'Location_Center_X': flts[:, 1].astype(np.float64),
        'Location_Center_Y': flts[:, 0].astype(np.float64),
    }
    for c, ch in enumerate(channels):
        o = 2 + 5 * c
        d[f'Intensity_IntegratedIntensity_{ch}'] = flts[:, o].astype(np.float64)
        d[f'Intensity_MeanIntensity_{ch}'] = flts[:, o + 1].astype(np.float64)
        d[f'Intensity_StdIntensity_{ch}'] = flts[:, o + 2].astype(np.float64)
        d[f'Intensity_MinIntensity_{ch}'] = flts[:, o + 3].astype(np.float64)
        d[f'Intensity_MaxIntensity_{ch}'] = flts[:, o + 4].astype(np.float64)
    return pd.DataFrame(d, columns=object_columns(channels))


def run(data_file, output, image_dir=None, batch=8, raw_intensities=False):
    import torch
    from .. import ops
    df = pd.read_csv(data_file)
    channels, illum_ch, objects = discover(df)
    if not channels or not objects:
        raise ValueError("LoadData CSV needs FileName_<channel> and Objects_FileName_<Object> columns")
    scale = 1.0 if raw_intensities else 1.0 / 65535.0
    os.makedirs(output, exist_ok=True)
    illum_dev = None
    if len(illum_ch) == len(channels):                         # plate-constant: load once
        first = df.iloc[0]
        fn = [np.load(_path(first, '', f'Illum{c}', image_dir)) for c in channels]
        illum_dev = torch.from_numpy(np.stack(fn).astype(np.float32)).cuda()
    image_rows = []
    frames = {o: [] for o in objects}

    def flush(pending):
        if not pending:
            return
        raw = torch.stack([p[1] for p in pending])[:, :, None].contiguous()               # [F][C][1][H][W], device
        ill = illum_dev if illum_dev is not None and tuple(illum_dev.shape[1:]) == tuple(raw.shape[3:]) else None
        for o in objects:
            labs = np.stack([p[2][o] for p in pending]).astype(np.int32)
            n_max = max(int(labs.max()), 1)
            res = ops.field_fused(raw, ill, torch.from_numpy(labs).cuda(), bin=1, intensity_scale=scale,
                                  n_max=n_max, want_maxproj=False, want_binned=False)
            n_obj = res["n_objects"].cpu().numpy()
            ints, flts = res["ints"].cpu().numpy(), res["flts"].cpu().numpy()
            for k, (image_number, _, _) in enumerate(pending):
                n = int(n_obj[k])
                frames[o].append(rows_to_frame(image_number, ints[k, :n], flts[k, :n], channels))
                image_rows[image_number - 1][f'Count_{o}'] = n
        pending.clear()

    pending, shape = [], None
    for i, (_, row) in enumerate(df.iterrows()):
        image_number = i + 1
        meta = {'ImageNumber': image_number}
        for col in df.columns:
            if col.startswith('Metadata_') or col.startswith('FileName_') or col.startswith('PathName_'):
                meta[col] = row[col]
        image_rows.append(meta)
        blobs = []
        for c in channels:
            with open(_path(row, '', c, image_dir), 'rb') as fh:
                blobs.append(fh.read())
        try:
            planes = tiffio.load_planes(blobs)                 # [C][H][W] on the device (TIFF strips decoded there, K7)
        except ValueError as e:
            raise ValueError("16-bit images of one shape expected (row %d): %s" % (image_number, e))
        labs = {o: tiffio.read(_path(row, 'Objects_', o, image_dir)) for o in objects}
        if shape is not None and tuple(planes.shape) != shape:
            flush(pending)
        shape = tuple(planes.shape)
        pending.append((image_number, planes, labs))
        if len(pending) == batch:
            flush(pending)
    flush(pending)
    image_df = pd.DataFrame(image_rows)
    for o in objects:
        image_df[f'Count_{o}'] = image_df[f'Count_{o}'].astype(np.int64)
        out = pd.concat(frames[o], ignore_index=True) if frames[o] else pd.DataFrame(columns=object_columns(channels))
        out.to_csv(os.path.join(output, f'{o}.csv'), index=False)
    image_df.to_csv(os.path.join(output, 'Image.csv'), index=False)
    return image_df, {o: os.path.join(output, f'{o}.csv') for o in objects}


def main(argv=None):
    a = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(message)s')
    run(a.data_file, a.output, a.image_dir, a.batch, a.raw_intensities)


if __name__ == '__main__':
    main()
