"""Helpers shared by the -m gpu parity tests (CUDA path through the C-ABI vs oracle/)."""
import numpy as np
import pytest


def require_gpu():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def dev(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def host(t):
    return t.detach().cpu().numpy()


def rel_err(got, ref, floor=0.0):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return np.abs(got - ref) / np.maximum(np.abs(ref), floor if floor > 0 else np.finfo(np.float64).tiny)
