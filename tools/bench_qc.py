"""K6 / A6 / A7 timing: the QC metrics of Illumination_QC_mult.py (radial power spectrum slope +
PercentMaximal) for one 2160^2 channel on the GPU path of the drop-in script, next to the
reference arithmetic (oracle/qc.py = the reference's own functions restated) on one host core."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from image_processing_suite_b200 import synth
from image_processing_suite_b200.scripts import Illumination_QC_mult as qc
from oracle import qc as o_qc

labs = synth.make_labels(2160, 2160, 2000, seed=3)
raw = synth.field_numpy(labs, c=5, z=1, seed=3)[:, 0]                 # [5][2160][2160] uint16
ill = [a for a in synth.make_illum(5, 2160, 2160, seed=0).astype(np.float64)]   # the script's illum_cache: one array per channel


def gpu_channel(c):
    corrected, pct = qc._corrected_and_pct(raw[c], ill[c])
    radii, _, powersum = qc.rps(corrected)
    return qc._slope(radii, powersum), pct


for c in range(5):          # first use uploads the channel's illumination function (once per plate)
    gpu_channel(c)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    got = [gpu_channel(c) for c in range(5)]
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / 20

t0 = time.perf_counter()
ref = []
for c in range(2):
    x = raw[c].astype(float) / ill[c]
    r = o_qc.qc_metrics(x, "ch")
    ref.append((r["ImageQuality_PowerLogLogSlope_ch"], r["ImageQuality_PercentMaximal_ch"]))
t_cpu = (time.perf_counter() - t0) / 2
for (gs, gp), (rs, rp) in zip(got, ref):
    assert abs(gs - rs) <= 1e-9 * max(1.0, abs(rs)) and gp == rp, (gs, rs, gp, rp)
print(json.dumps({"step": "QC metrics of one 2160^2 channel (divide, PercentMaximal, FFT, ring sums, slope), host image in",
                  "gpu_ms_per_channel": t_gpu * 1e3, "cpu_reference_ms_per_channel_1core": t_cpu * 1e3,
                  "channels_per_s_gpu": 1 / t_gpu, "channels_per_s_cpu_1core": 1 / t_cpu, "results_match": True}))
