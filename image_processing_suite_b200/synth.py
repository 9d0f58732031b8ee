"""Seeded synthetic inputs for the hot path.

Fields are 5-channel uint16 z-stacks with Cellpose-like label masks (non-overlapping
rotated ellipses, semi-axes 14..24 px, ~45 % foreground at 2000 cells / 2160^2), a smooth
illumination function in [1, 1.5], and a few saturated pixels.  ``field_numpy`` is the
host generator used by the tests; ``field_torch`` builds the same kind of field on a
device for the benchmark ring (the CPU baseline then works on a D2H copy of the very
same bytes, so both arms always see identical inputs).
"""
import numpy as np


def make_labels(h, w, n_cells, seed=0, amin=14, amax=24, max_tries=None, dtype=np.int32):
    """Label mask [h][w]: 0 background, 1..N contiguous (N <= n_cells if space runs out)."""
    rng = np.random.default_rng(seed)
    lab = np.zeros((h, w), dtype)
    placed = 0
    tries = 0
    max_tries = max_tries or 60 * n_cells + 1000
    while placed < n_cells and tries < max_tries:
        tries += 1
        a = int(rng.integers(amin, amax + 1))
        b = int(rng.integers(amin, amax + 1))
        th = rng.uniform(0, np.pi)
        r = max(a, b)
        cy = int(rng.integers(0, h))
        cx = int(rng.integers(0, w))
        y0, y1 = max(cy - r, 0), min(cy + r + 1, h)
        x0, x1 = max(cx - r, 0), min(cx + r + 1, w)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dy, dx = yy - cy, xx - cx
        u = (dx * np.cos(th) + dy * np.sin(th)) / a
        v = (-dx * np.sin(th) + dy * np.cos(th)) / b
        m = (u * u + v * v) <= 1.0
        sub = lab[y0:y1, x0:x1]
        if not m.any() or (sub[m] != 0).any():
            continue
        placed += 1
        sub[m] = placed
    return lab


def make_illum(c, h, w, seed=0, dtype=np.float32):
    """Smooth per-channel illumination function in [1, 1.5]."""
    rng = np.random.default_rng(seed + 7919)
    yy = (np.arange(h, dtype=np.float64)[:, None] - h / 2.0) / max(h, 1)
    xx = (np.arange(w, dtype=np.float64)[None, :] - w / 2.0) / max(w, 1)
    out = np.empty((c, h, w), dtype)
    for k in range(c):
        oy, ox = rng.uniform(-0.1, 0.1, 2)
        r2 = (yy - oy) ** 2 + (xx - ox) ** 2
        f = 1.0 / (1.0 + 2.0 * r2)                      # vignette, (0, 1]
        f = (f - f.min()) / max(f.max() - f.min(), 1e-12)
        out[k] = (1.0 + 0.5 * f).astype(dtype)
    return out


def field_numpy(labels, c=5, z=3, seed=0, saturate_frac=1e-4):
    """raw[c][z][h][w] uint16 for one field with the given label mask."""
    rng = np.random.default_rng(seed + 104729)
    h, w = labels.shape
    n = int(labels.max()) if labels.size else 0
    amp = np.r_[0.0, rng.uniform(500.0, 8000.0, n)]
    gain = rng.uniform(0.5, 2.0, c)
    z0 = rng.uniform(0, max(z - 1, 0) + 1e-9)
    yy = (np.arange(h)[:, None] - h / 2.0) / max(h, 1)
    xx = (np.arange(w)[None, :] - w / 2.0) / max(w, 1)
    vign = 1.0 / (1.0 + 0.5 * (yy * yy + xx * xx) * 4.0)
    cell = amp[labels]
    raw = np.empty((c, z, h, w), np.uint16)
    for k in range(c):
        for p in range(z):
            att = np.exp(-((p - z0) / 1.5) ** 2)
            tex = 1.0 + 0.15 * rng.standard_normal((h, w))
            img = (rng.normal(300.0, 30.0, (h, w)) + cell * gain[k] * att * tex) * vign
            raw[k, p] = np.clip(np.rint(img), 0, 65535).astype(np.uint16)
    n_sat = int(round(saturate_frac * h * w))
    if n_sat:
        ys = rng.integers(0, h, n_sat)
        xs = rng.integers(0, w, n_sat)
        raw[:, rng.integers(0, z), ys, xs] = 65535
    return raw


def field_torch(labels_dev, c=5, z=3, seed=0, saturate_frac=1e-4):
    """Device generator: raw[c][z][h][w] uint16 tensor on ``labels_dev.device``."""
    import torch
    dev = labels_dev.device
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) + 104729)
    h, w = labels_dev.shape
    n = int(labels_dev.max().item()) if labels_dev.numel() else 0
    amp = torch.empty(n + 1, device=dev).uniform_(500.0, 8000.0, generator=g)
    amp[0] = 0.0
    cell = amp[labels_dev.long()]
    gain = torch.empty(c, device=dev).uniform_(0.5, 2.0, generator=g)
    z0 = float(torch.empty(1, device=dev).uniform_(0, max(z - 1, 0) + 1e-9, generator=g).item())
    yy = (torch.arange(h, device=dev, dtype=torch.float32)[:, None] - h / 2.0) / max(h, 1)
    xx = (torch.arange(w, device=dev, dtype=torch.float32)[None, :] - w / 2.0) / max(w, 1)
    vign = 1.0 / (1.0 + 2.0 * (yy * yy + xx * xx))
    raw = torch.empty((c, z, h, w), dtype=torch.uint16, device=dev)
    for k in range(c):
        for p in range(z):
            att = float(np.exp(-((p - z0) / 1.5) ** 2))
            tex = 1.0 + 0.15 * torch.randn((h, w), device=dev, generator=g)
            bg = 300.0 + 30.0 * torch.randn((h, w), device=dev, generator=g)
            img = (bg + cell * (gain[k] * att) * tex) * vign
            raw[k, p] = img.round_().clamp_(0, 65535).to(torch.int32).to(torch.uint16)
    n_sat = int(round(saturate_frac * h * w))
    if n_sat:
        ys = torch.randint(0, h, (n_sat,), device=dev, generator=g)
        xs = torch.randint(0, w, (n_sat,), device=dev, generator=g)
        view = raw.view(torch.int16)
        view[:, 0, ys, xs] = -1                        # 0xFFFF
    return raw
