"""K6 / A6 / A7 timing: the QC metrics of Illumination_QC_mult.py (radial power spectrum slope +
PercentMaximal) for one 2160^2 channel on the GPU path of the drop-in script, next to the
reference's NumPy / SciPy arithmetic (Illumination_QC_mult.py:31-125, restated below for timing)
on one host core."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from image_processing_suite_b200 import synth
from image_processing_suite_b200.scripts import Illumination_QC_mult as qc



def cpu_qc(x):
    """rps + slope + PercentMaximal as the reference computes them (:31-125), NumPy / SciPy on the host."""
    import scipy.fft
    import scipy.ndimage
    import scipy.stats
    h, w = x.shape
    y = x / np.median(np.abs(x - x.mean())) if np.ptp(x) > 0 else x
    spec = np.abs(scipy.fft.fft2(y - y.mean()))
    di = np.minimum(np.arange(h), h - 1 - np.arange(h))
    dj = np.minimum(np.arange(w), w - 1 - np.arange(w))
    rings = np.floor(np.sqrt(di[:, None] ** 2 + dj[None, :] ** 2)).astype(np.int64) + 1
    labels = np.arange(2, int(np.floor(min(h, w) / 8.0)))
    power = np.asarray(scipy.ndimage.sum(spec ** 2, rings, labels))
    ok = power > 0
    slope = scipy.stats.linregress(np.log(labels[ok]), np.log(power[ok]))[0] if ok.sum() > 2 else 0.0
    return slope, 100.0 * float(np.count_nonzero(x == x.max())) / x.size


labs = synth.make_labels(2160, 2160, 2000, seed=3)
raw = synth.field_numpy(labs, c=5, z=1, seed=3)[:, 0]                 # [5][2160][2160] uint16
ill = [a for a in synth.make_illum(5, 2160, 2160, seed=0).astype(np.float64)]   # the script's illum_cache: one array per channel


def gpu_channel(c):
    return qc._channel_metrics(raw[c], ill[c])      # host uint16 image in, (slope, PercentMaximal) out, one sync


for c in range(5):          # first use uploads the channel's illumination function (once per plate)
    gpu_channel(c)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    got = [gpu_channel(c) for c in range(5)]
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / 20

# the same with the images already on the device (what process_site sees: TIFF strips are decoded there)
raw_dev = [torch.from_numpy(raw[c]).cuda() for c in range(5)]
for c in range(5):
    qc._channel_metrics(raw_dev[c], ill[c])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    got_dev = [qc._channel_metrics(raw_dev[c], ill[c]) for c in range(5)]
torch.cuda.synchronize()
t_dev = (time.perf_counter() - t0) / 20
assert got_dev == got

t0 = time.perf_counter()
ref = []
for c in range(2):
    x = raw[c].astype(float) / ill[c]
    ref.append(cpu_qc(x))
t_cpu = (time.perf_counter() - t0) / 2
for (gs, gp), (rs, rp) in zip(got, ref):
    assert abs(gs - rs) <= 1e-9 * max(1.0, abs(rs)) and gp == rp, (gs, rs, gp, rp)
print(json.dumps({"step": "QC metrics of one 2160^2 channel (divide, PercentMaximal, FFT, ring sums, slope), host image in",
                  "gpu_ms_per_channel": t_gpu * 1e3, "gpu_ms_per_channel_device_image_in": t_dev * 1e3, "cpu_reference_ms_per_channel_1core": t_cpu * 1e3,
                  "channels_per_s_gpu": 1 / t_gpu, "channels_per_s_cpu_1core": 1 / t_cpu, "results_match": True}))
