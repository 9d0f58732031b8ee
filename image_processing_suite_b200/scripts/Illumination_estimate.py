"""Per-plate illumination-function estimation (north_star; no reference script -- the
reference loads functions produced elsewhere, Illumination_QC_mult.py:186-193).

Reads the LoadData CSV, streams every site's channel images through ips_illum_accumulate
(or keeps them resident for --mode median), and writes ``{illum_path}/{ch}_illum.npy``
(float32, same H x W as the images) -- the format Illumination_QC_mult.py and
Cellpose_GPU_s3fs.py:56 load.

Several plates / timepoints: give one LoadData CSV per plate-timepoint; the functions of CSV
``<name>.csv`` go to ``{illum_path}/<name>/``.  Under ``torchrun`` (RANK / WORLD_SIZE, or --rank /
--world) the (plate, channel) units are dealt over the ranks (plate.shard_units): the estimation
reduces across the fields of a plate, so this -- not the field -- is its unit of sharding, and
there is no communication at all.
"""
import argparse
import logging
import os

import numpy as np
import pandas as pd

from . import tiffio


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Estimate per-channel illumination functions for one plate")
    p.add_argument('--load-data', type=str, nargs='+', required=True,
                   help="LoadData CSV with FileName_{ch} columns; several = one per plate / timepoint")
    p.add_argument('--data-path', type=str, required=True, help="Base path for image files")
    p.add_argument('--illum-path', type=str, required=True, help="Output folder for {ch}_illum.npy")
    p.add_argument('--channels', nargs='+', required=True)
    p.add_argument('--filter-size', type=float, default=200.0, help="Gaussian FWHM in pixels (sigma = size / 2.35)")
    p.add_argument('--robust-frac', type=float, default=0.02)
    p.add_argument('--mode', choices=['mean', 'median'], default='mean')
    p.add_argument('--batch', type=int, default=16, help="sites per device batch")
    p.add_argument('--rank', type=int, default=int(os.environ.get("RANK", "0")))
    p.add_argument('--world', type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    return p.parse_args(argv)


def estimate(load_data, data_path, channels, filter_size=200.0, robust_frac=0.02, mode='mean', batch=16):
    """Returns {channel: float32 [H][W]}."""
    import torch
    from .. import ops
    df = pd.read_csv(load_data)
    cols = [f'FileName_{c}' for c in channels]
    est, stack, buf = None, [], []

    def flush():
        nonlocal est
        if not buf:
            return
        flat = tiffio.load_planes([b for site in buf for b in site])      # TIFF strips decoded on the GPU (K7)
        dev = flat.reshape(len(buf), len(cols), flat.shape[1], flat.shape[2])   # [F][C][H][W]
        if mode == 'mean':
            if est is None:
                est = ops.IllumEstimator(dev.shape[1], dev.shape[2], dev.shape[3])
            est.add(dev)
        else:
            stack.append(dev)
        buf.clear()

    for _, row in df.iterrows():
        site = []
        for c in cols:
            with open(os.path.join(data_path, row[c]), "rb") as fh:
                site.append(fh.read())
        buf.append(site)
        if len(buf) == batch:
            flush()
    flush()
    sigma = float(filter_size) / 2.35
    if mode == 'mean':
        if est is None:
            raise ValueError("no sites in %s" % load_data)
        out = est.finalize(sigma, robust_frac)
    else:
        raw = ops.illum_median(torch.cat(stack).contiguous())
        out = ops.illum_smooth_rescale(raw, sigma, robust_frac)
    host = out.cpu().numpy()
    return {c: host[i] for i, c in enumerate(channels)}


def main(argv=None):
    from ..plate import shard_units
    a = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(message)s')
    if a.world > 1:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", a.rank)) % max(torch.cuda.device_count(), 1))
    mine = shard_units(len(a.load_data), a.channels, a.rank, a.world)
    written = {}
    for p, csv_path in enumerate(a.load_data):
        chans = [c for (q, c) in mine if q == p]
        if not chans:
            continue
        funcs = estimate(csv_path, a.data_path, chans, a.filter_size, a.robust_frac, a.mode, a.batch)
        out_dir = a.illum_path if len(a.load_data) == 1 else os.path.join(
            a.illum_path, os.path.splitext(os.path.basename(csv_path))[0])
        os.makedirs(out_dir, exist_ok=True)
        for c, f in funcs.items():
            np.save(os.path.join(out_dir, f"{c}_illum.npy"), f)
            logging.info(f"[rank {a.rank}/{a.world}] wrote {out_dir}/{c}_illum.npy {f.shape} min {f.min():.4f} max {f.max():.4f}")
            written[(p, c)] = f
    return written if len(a.load_data) > 1 else {c: f for (_, c), f in written.items()}


if __name__ == '__main__':
    main()
