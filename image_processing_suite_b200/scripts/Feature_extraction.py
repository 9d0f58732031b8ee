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
    p.add_argument('--threads', type=int, default=8, help="reader threads staging the next batches")
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


class _ObjectCsv:
    """One <Object>.csv written batch by batch: the device rows of a batch arrive as one dense float32
    table (plate.pack_rows), become Arrow columns (integers as int64, floats as float64 of the float32
    values) and are appended by Arrow's multi-threaded CSV writer on a writer thread."""

    def __init__(self, path, channels, writers):
        import pyarrow as pa
        self._pa = pa
        self.path, self.channels, self._writers = path, channels, writers
        self._cols = object_columns(channels)
        self._file = open(path, "wb")
        self._file.write((",".join(self._cols) + "\n").encode())
        self._last = None

    def append(self, rows):
        """rows float32 [n][10 + 5C]: image number, batch index, label, area, y0, x0, y1, x1, cy, cx, per channel 5."""
        if rows.shape[0] == 0:
            return
        pa = self._pa
        i64 = lambda c: pa.array(rows[:, c].astype(np.int64))
        f64 = lambda c: pa.array(rows[:, c].astype(np.float64))
        cols = [i64(0), i64(2), i64(3), i64(5), i64(4), i64(7), i64(6), f64(9), f64(8), f64(9), f64(8)]
        for c in range(len(self.channels)):
            cols += [f64(10 + 5 * c + k) for k in range(5)]
        table = pa.Table.from_arrays(cols, names=self._cols)
        prev = self._last

        def write():
            import pyarrow.csv as pacsv
            sink = pa.BufferOutputStream()         # formatted on this thread, in parallel with the other batches
            pacsv.write_csv(table, sink, write_options=pacsv.WriteOptions(include_header=False, quoting_style="none"))
            text = sink.getvalue()
            if prev is not None:
                prev.result()                      # ... and appended to the file in batch order
            self._file.write(memoryview(text))
        self._last = self._writers.submit(write)

    def close(self):
        if self._last is not None:
            self._last.result()
        self._file.close()


def run(data_file, output, image_dir=None, batch=8, raw_intensities=False, threads=8):
    """LoadData CSV -> Image.csv + one <Object>.csv per label set.  Sites are staged ``batch`` at a time
    by reader threads (scripts/batchio.py: image and label files read into page-locked memory, one
    host->device copy, TIFF strips decoded on the device), measured by ONE ips_field_fused launch per
    label set and batch (label masks stay uint16), and the rows leave through one dense device->host
    copy and Arrow's CSV writer while the next batch is read."""
    import torch
    from .. import ops, plate
    from . import batchio
    df = pd.read_csv(data_file)
    channels, illum_ch, objects = discover(df)
    if not channels or not objects:
        raise ValueError("LoadData CSV needs FileName_<channel> and Objects_FileName_<Object> columns")
    scale = 1.0 if raw_intensities else 1.0 / 65535.0
    os.makedirs(output, exist_ok=True)
    illum_dev = illum_rcp = None
    if len(illum_ch) == len(channels):                         # plate-constant: load once
        first = df.iloc[0]
        fn = [np.load(_path(first, '', f'Illum{c}', image_dir)) for c in channels]
        illum_dev = torch.from_numpy(np.stack(fn).astype(np.float32)).cuda()
        illum_rcp = ops.illum_reciprocal(illum_dev)            # once per plate: the divide becomes a multiply
    C_, O_ = len(channels), len(objects)
    rows_meta = []
    for i, (_, row) in enumerate(df.iterrows()):
        meta = {'ImageNumber': i + 1}
        for col in df.columns:
            if col.startswith('Metadata_') or col.startswith('FileName_') or col.startswith('PathName_'):
                meta[col] = row[col]
        rows_meta.append(meta)
    site_files = [[_path(row, '', c, image_dir) for c in channels] + [_path(row, 'Objects_', o, image_dir) for o in objects]
                  for _, row in df.iterrows()]

    def batches():
        for i in range(0, len(site_files), batch):
            yield (i, min(i + batch, len(site_files))), [p for site in site_files[i:i + batch] for p in site]

    writers = batchio.Writers(threads=max(2, threads))
    csvs = {o: _ObjectCsv(os.path.join(output, f'{o}.csv'), channels, writers) for o in objects}
    counts = {o: np.zeros(len(site_files), np.int64) for o in objects}

    def measure(i0, planes):
        """planes uint16 [B][C + O][H][W] on the device -> rows of every label set appended."""
        B = planes.shape[0]
        raw = planes[:, :C_].unsqueeze(2).contiguous()                                  # [B][C][1][H][W]
        H, W = planes.shape[2:]
        use_illum = illum_dev is not None and tuple(illum_dev.shape[1:]) == (H, W)
        numbers = torch.arange(i0 + 1, i0 + 1 + B, device=planes.device, dtype=torch.int32)  # ImageNumber rides in the row's field column
        for k, o in enumerate(objects):
            labs = planes[:, C_ + k].contiguous()                                        # uint16 masks as they are
            n_max = max(int(labs.to(torch.int32).max().item()), 1)
            res = ops.field_fused(raw, illum_dev if use_illum else None, labs, bin=1, intensity_scale=scale, n_max=n_max,
                                  want_maxproj=False, want_binned=False, illum_rcp=illum_rcp if use_illum else None)
            n_obj = res["n_objects"]
            if bool((n_obj < 0).any()):
                raise ValueError("label mask with values above its own maximum")          # cannot happen
            rows, total = plate.pack_rows(res["ints"], res["flts"], n_obj, numbers, field_base=0)
            host_n = n_obj.cpu().numpy()
            n = int(host_n.sum())
            host_rows = torch.empty((n, rows.shape[1]), dtype=torch.float32, pin_memory=True)
            host_rows.copy_(rows[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            csvs[o].append(host_rows.numpy())
            counts[o][i0:i0 + B] = host_n

    loader = batchio.BatchLoader(batches(), threads=threads, depth=3)
    import time
    trace = batchio.Trace()
    it = iter(loader)
    while True:
        t0 = time.perf_counter()
        staged = next(it, None)
        trace.add("wait for staged files", t0)
        if staged is None:
            break
        i0, i1 = staged.tag
        try:
            if not staged.ok():
                raise ValueError("a file of the batch is not a TIFF the device codec reads")
            t0 = time.perf_counter()
            src = batchio.to_device(staged)
            planes = tiffio.decode_staged(src, staged.infos, staged.bases)
            trace.add("copy + decode", t0)
            t0 = time.perf_counter()
            measure(i0, planes.view(i1 - i0, C_ + O_, planes.shape[1], planes.shape[2]))
            trace.add("measure + rows to host", t0)
        except (ValueError, tiffio.Unsupported) as why:
            # 8-bit / tiled / deflate images or masks, mixed shapes: the host decoder, site by site
            logger.info("sites %d-%d use the host image decoder (%s)", i0 + 1, i1, why)
            for i in range(i0, i1):
                blobs = [open(p, 'rb').read() for p in site_files[i][:C_]]
                try:
                    img = tiffio.load_planes(blobs)
                except ValueError as e:
                    raise ValueError("16-bit images of one shape expected (row %d): %s" % (i + 1, e))
                labs = [torch.from_numpy(np.ascontiguousarray(tiffio.read(p)).astype(np.uint16)).cuda() for p in site_files[i][C_:]]
                measure(i, torch.cat([img, torch.stack(labs)])[None])
        finally:
            loader.release(staged)
    t0 = time.perf_counter()
    for o in objects:
        csvs[o].close()
    writers.close()
    trace.add("wait for writers", t0)
    trace.report(logger, "Feature_extraction")
    if writers.errors:
        raise IOError("writing the object tables failed: %s" % writers.errors[0])
    image_df = pd.DataFrame(rows_meta)
    for o in objects:
        image_df[f'Count_{o}'] = counts[o]
    image_df.to_csv(os.path.join(output, 'Image.csv'), index=False)
    return image_df, {o: os.path.join(output, f'{o}.csv') for o in objects}


def main(argv=None):
    a = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(message)s')
    run(a.data_file, a.output, a.image_dir, a.batch, a.raw_intensities, a.threads)


if __name__ == '__main__':
    main()
