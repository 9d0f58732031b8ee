"""Drop-in for Illumination_QC_mult.py: same flags, same functions, same output columns.

Per site x channel: read the TIFF, divide by the pre-loaded illumination function
(Illumination_QC_mult.py:145-150), ImageQuality_PowerLogLogSlope (rps, :31-70, :104-116) and
ImageQuality_PercentMaximal (:73-95).  On the device, one stream, one device -> host read per
channel: float64 divide + mean + exact radix-select median (ips_rps_prepare), real FFT (cuFFT via
torch.fft.rfft2 is the FFT library, as scipy.fftpack is in the reference), ring-keyed sums over
the Hermitian half (ips_ring_sums_half), log-log least squares (ips_loglog_slope); PercentMaximal
is the float64 side reduction of ips_preprocess_fused.
Error convention unchanged: process_site never raises; failures become QC_Error_{ch} strings.
"""
import argparse
import concurrent.futures
import logging
import os
import threading

import numpy as np
import pandas as pd

from . import tiffio


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="CellProfiler-Matched Image QC (Production)")
    parser.add_argument('--load-data', type=str, required=True, help="Path to input CSV (LoadData format)")
    parser.add_argument('--data-path', type=str, required=True, help="Base path for image files")
    parser.add_argument('--illum-path', type=str, default=None, help="Folder containing .npy illumination functions")
    parser.add_argument('--channels', nargs='+', required=True, help="List of channel names (e.g. CL488 CL568)")
    parser.add_argument('--output', type=str, default='QC_Results.csv', help="Path for output CSV")
    parser.add_argument('--threads', type=int, default=24, help="Number of threads for parallel processing")
    return parser.parse_args(argv)


_streams = threading.local()


def _stream():
    """One CUDA stream per worker thread (process_site is called from a thread pool, :212)."""
    import torch
    s = getattr(_streams, "s", None)
    if s is None:
        s = _streams.s = torch.cuda.Stream()
    return s


def _to_device_f64(image):
    import torch
    if isinstance(image, torch.Tensor):
        return image.to(device="cuda", dtype=torch.float64)
    return torch.from_numpy(np.ascontiguousarray(image, dtype=np.float64)).cuda()


def rps(img):
    """Radial power spectrum ring sums, Illumination_QC_mult.py:31-70.
    Returns (labels, magsum, powersum) as NumPy arrays, or the reference's degenerate
    ``[2], [0], [0]`` lists when min(H, W) < 24 (no ring to sum, :70).
    On the device end to end: exact radix-select median, real FFT, ring sums over the Hermitian
    half (ops.rps_spectrum)."""
    from .. import ops
    x = _to_device_f64(img)
    assert x.dim() == 2
    H, W = x.shape
    labels = np.arange(2, np.floor(min(H, W) / 8.0)).astype(int)
    if len(labels) == 0:
        return [2], [0], [0]
    mag, pw = ops.rps_spectrum(x.contiguous())
    return labels, mag.cpu().numpy(), pw.cpu().numpy()


def calculate_saturation_cp_exact(image, mask=None):
    """PercentMaximal, Illumination_QC_mult.py:73-95: 100 * #(x == max x) / #x."""
    import torch
    x = _to_device_f64(image)
    if mask is not None:
        x = x[torch.as_tensor(mask, device=x.device)]
    n = x.numel()
    if n == 0:
        return 0.0
    return 100.0 * float((x == x.max()).sum().item()) / float(n)


def _slope(radii, powersum):
    import scipy.stats
    valid = powersum > 0                                     # TypeError for the list case -> NaN
    if np.sum(valid) > 2:
        return scipy.stats.linregress(np.log(radii[valid]), np.log(powersum[valid]))[0]
    return 0.0


def calculate_qc_metrics(image, channel_name):
    """Both metrics of one channel, :98-125 (a failing metric becomes NaN)."""
    results = {}
    try:
        radii, _, powersum = rps(image)
        results[f'ImageQuality_PowerLogLogSlope_{channel_name}'] = _slope(radii, powersum)
    except Exception:
        results[f'ImageQuality_PowerLogLogSlope_{channel_name}'] = np.nan
    try:
        results[f'ImageQuality_PercentMaximal_{channel_name}'] = calculate_saturation_cp_exact(image)
    except Exception:
        results[f'ImageQuality_PercentMaximal_{channel_name}'] = np.nan
    return results


_illum_dev = {}
_illum_lock = threading.Lock()


def _illum_on_device(illum):
    """(float64 tensor, float32 tensor or None) of one illumination function, uploaded once per
    array (the cache of load_illum_cache is shared by every site, Illumination_QC_mult.py:196-199).
    The float32 copy exists only when the function is float32-exact, which is when the fused
    kernel's divide-side PercentMaximal applies."""
    import torch
    illum = np.asarray(illum)
    key = (illum.__array_interface__["data"][0], illum.shape, illum.strides, illum.dtype.str)
    with _illum_lock:
        hit = _illum_dev.get(key)
        if hit is not None:
            return hit[1], hit[2]
    ill64 = np.ascontiguousarray(illum, dtype=np.float64)
    ill32 = ill64.astype(np.float32)
    d64 = torch.from_numpy(ill64).cuda()
    d32 = torch.from_numpy(ill32[None]).cuda() if np.array_equal(ill32.astype(np.float64), ill64) else None
    with _illum_lock:
        if len(_illum_dev) > 64:
            _illum_dev.clear()
        _illum_dev[key] = (illum, d64, d32)          # keeps the memory alive, so the key stays unique
    return d64, d32


def _channel_metrics(img_u16, illum):
    """(PowerLogLogSlope, PercentMaximal) of one uint16 channel image on the device, as device /
    host scalars resolved with ONE synchronisation: the float64 divide, the exact median, the
    real FFT, the ring sums and the log-log regression all stay on the stream (ops.rps_spectrum,
    ops.loglog_slope); PercentMaximal is the float64 side reduction of the fused kernel when the
    function is float32-exact, else an equality count on the float64 corrected image."""
    import torch
    from .. import ops
    raw = img_u16 if isinstance(img_u16, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img_u16)).cuda()
    raw = raw.contiguous()
    d64 = d32 = None
    if illum is not None and tuple(img_u16.shape) == tuple(illum.shape):      # else silently uncorrected (:148-153)
        d64, d32 = _illum_on_device(illum)
    need_corrected = d64 is not None and d32 is None
    out = ops.rps_spectrum(raw, d64, want_corrected=need_corrected)
    pw = out[1]
    slope = ops.loglog_slope(pw) if pw.numel() else None             # no ring at all: the reference's NaN (:70, :115)
    nan = torch.full((), float("nan"), dtype=torch.float64, device=raw.device)
    if need_corrected:
        corrected = out[2]
        count = (corrected == corrected.max()).sum().to(torch.float64)
        got = torch.stack([slope[0] if slope is not None else nan, count]).tolist()      # the one device -> host read
        # the reference's own expression, evaluated on the host (:93; a device divide by a scalar may
        # be a multiply by its reciprocal)
        return got[0], 100.0 * float(got[1]) / float(corrected.numel())
    r = ops.preprocess_fused(raw[None, None, None], d32, bin=1, want_maxproj=False, want_binned=False,
                             want_pct_maximal=True)
    got = torch.stack([slope[0] if slope is not None else nan, r["pct_maximal"][0, 0]]).tolist()
    return got[0], got[1]


def process_site(site_data):
    """One CSV row: (index, paths, channels, illum_cache) -> (index, {column: value}).
    Never raises; see Illumination_QC_mult.py:131-162."""
    import torch
    index, paths, channels, illum_cache = site_data
    site_results = {}
    with torch.cuda.stream(_stream()):
        for i, (path, ch_name) in enumerate(zip(paths, channels)):
            try:
                if not os.path.exists(path):
                    site_results[f"QC_Error_{ch_name}"] = "File Not Found"
                    continue
                with open(path, "rb") as fh:
                    data = fh.read()
                try:
                    img = tiffio.decode_to_device([data])[0]         # strips decoded on the GPU (K7)
                except tiffio.Unsupported:
                    img = tiffio.decode(data)
                illum = illum_cache[i] if illum_cache and illum_cache[i] is not None else None
                if isinstance(img, torch.Tensor) or (img.dtype == np.uint16 and img.ndim == 2):
                    try:
                        slope, pct = _channel_metrics(img, illum)
                    except Exception:
                        slope = pct = np.nan
                    site_results[f'ImageQuality_PowerLogLogSlope_{ch_name}'] = slope
                    site_results[f'ImageQuality_PercentMaximal_{ch_name}'] = pct
                else:
                    x = img.astype(float)
                    if illum is not None and x.shape == illum.shape:
                        x = x / illum
                    site_results.update(calculate_qc_metrics(x, ch_name))
            except Exception as e:
                site_results[f"QC_Error_{ch_name}"] = str(e)
        torch.cuda.current_stream().synchronize()
    return index, site_results


ILLUM_FILE_PATTERNS = ("{ch}_illum.npy", "Illum{ch}.npy")      # the two names the reference accepts (:186-187)


def load_illum_cache(illum_path, channels):
    """One entry per channel: the first of ``{ch}_illum.npy`` / ``Illum{ch}.npy`` found under
    ``illum_path``, else None (the channel then stays uncorrected, :148, :199)."""
    def find(ch):
        for pattern in ILLUM_FILE_PATTERNS:
            name = pattern.format(ch=ch)
            full = os.path.join(illum_path, name)
            if os.path.exists(full):
                logging.info("  Loaded %s", name)
                return np.load(full)
        logging.warning("  Warning: No illumination file found for %s", ch)
        return None
    return [find(ch) if illum_path else None for ch in channels]


def run(load_data, data_path, channels, illum_path=None, output='QC_Results.csv', threads=24):
    """The script body: LoadData CSV in, the same table plus the ImageQuality_* / QC_Error_*
    columns out.  Stale QC columns of an earlier run are dropped first; sites fan out over a
    thread pool (one CUDA stream per worker) and are re-assembled by CSV index."""
    table = pd.read_csv(load_data)
    table = table[[c for c in table.columns if 'ImageQuality_' not in c and 'QC_Error' not in c]]
    illum = load_illum_cache(illum_path, channels)
    names = table[[f'FileName_{ch}' for ch in channels]]
    sites = [(idx, [os.path.join(data_path, n) for n in row], channels, illum)
             for idx, row in zip(names.index, names.itertuples(index=False, name=None))]
    logging.info("Starting processing on %d sites with %d threads...", len(sites), threads)
    with concurrent.futures.ThreadPoolExecutor(max_workers=threads) as pool:
        per_site = dict(pool.map(process_site, sites))
    qc = pd.DataFrame.from_dict(per_site, orient='index').sort_index()
    result = pd.concat([table, qc], axis=1)
    result.to_csv(output, index=False)
    logging.info("Done! Saved to %s", output)
    return result


def main(argv=None):
    a = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(message)s')
    return run(a.load_data, a.data_path, a.channels, a.illum_path, a.output, a.threads)


def bind_reference(path):
    """The other way to deploy: load the UNMODIFIED reference script from ``path`` and substitute
    the GPU functions for its arithmetic (process_site, rps, calculate_saturation_cp_exact,
    calculate_qc_metrics); its own ``main()`` / argparse / CSV handling then run as they are.
    Returns the bound module (call ``.main()``)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("Illumination_QC_mult_reference", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name in ("process_site", "rps", "calculate_saturation_cp_exact", "calculate_qc_metrics"):
        setattr(ref, name, globals()[name])
    return ref


if __name__ == '__main__':
    main()
