import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from image_processing_suite_b200 import ops, capi
from oracle import qc as o_qc
rng = np.random.default_rng(0)
H, W = 64, 64
yy, xx = np.mgrid[0:H, 0:W]
x = 2000.0 + 900.0 * np.sin(yy / 7.0) * np.cos(xx / 5.0) + rng.normal(0, 60.0, (H, W))
dx = torch.from_numpy(x).cuda()
z = torch.empty_like(dx)
ws = torch.zeros(int(capi.call("ips_rps_prepare_workspace_bytes", H * W)), dtype=torch.uint8, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
capi.call("ips_rps_prepare", None, p(dx), None, None, p(z), H * W, p(ws), ws.numel(), None)
torch.cuda.synchronize()
st = ws[:80].cpu().numpy()
print("state doubles", st[:32].view(np.float64), "u64", st[32:72].view(np.uint64), "ints", st[72:80].view(np.int32))
med = np.median(np.abs(x - x.mean()))
print("expect mean", x.mean(), "min", x.min(), "max", x.max(), "median", med, "key", np.float64(med).view(np.uint64))
ze = x / med - x.mean() / med
print("z max abs err", np.abs(z.cpu().numpy() - ze).max())
spec = torch.fft.rfft2(z); print("strides", spec.stride(), spec.is_contiguous()); spec = spec.contiguous()
print("spec err", np.abs(spec.cpu().numpy() - np.fft.rfft2(ze)).max())
labels, e_mag, e_pow = o_qc.radial_power_spectrum(x)
mag = torch.empty(len(labels), dtype=torch.float64, device="cuda"); pw = torch.empty_like(mag)
capi.call("ips_ring_sums_half", p(spec), p(mag), p(pw), len(labels), 1, H, W, None)
print(mag.cpu().numpy()[:6], e_mag[:6])
