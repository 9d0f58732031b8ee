"""K1 parity: fused z-max -> illum divide -> bin (CUDA, through the C-ABI) vs the oracle.

Bit-exact: max projection, integer bin sums, PercentMaximal.  fp32 vs float64: corrected
pixels and float bin sums within RTOL = 1e-5 (north_star's tolerance).
"""
import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import preprocess as o_pre
from oracle import qc as o_qc
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _fields(F, C, Z, H, W, seed, cells=12):
    labs, raws = [], []
    for f in range(F):
        lab = synth.make_labels(H, W, cells, seed=seed + f, amin=4, amax=9)
        labs.append(lab)
        raws.append(synth.field_numpy(lab, c=C, z=Z, seed=seed + f, saturate_frac=2e-3))
    return np.stack(labs), np.stack(raws)


@pytest.mark.parametrize("shape", [(2, 3, 3, 64, 96), (1, 5, 5, 48, 40), (3, 2, 1, 32, 64), (1, 1, 7, 40, 72)])
@pytest.mark.parametrize("bin", [1, 2, 4])
def test_preprocess_with_illum(shape, bin):
    require_gpu()
    from image_processing_suite_b200 import ops
    F, C, Z, H, W = shape
    _, raw = _fields(F, C, Z, H, W, seed=11)
    ill = synth.make_illum(C, H, W, seed=2)
    r = ops.preprocess_fused(dev(raw), dev(ill), bin=bin, want_corrected=True, want_pct_maximal=True)
    mp, corr, binned, pct = host(r["maxproj"]), host(r["corrected"]), host(r["binned"]), host(r["pct_maximal"])
    for f in range(F):
        emp, ecorr, ebin = o_pre.preprocess_field(raw[f], ill, bin)
        np.testing.assert_array_equal(mp[f], emp)
        np.testing.assert_allclose(corr[f], ecorr, rtol=RTOL, atol=0)
        np.testing.assert_allclose(binned[f], ebin, rtol=RTOL, atol=0)
        for c in range(C):
            assert pct[f, c] == o_qc.percent_maximal(o_pre.illum_correct(emp[c], ill[c].astype(np.float64)))


@pytest.mark.parametrize("shape", [(2, 3, 3, 64, 96), (1, 5, 2, 48, 40)])
@pytest.mark.parametrize("bin", [1, 2, 4])
def test_preprocess_integer_mode(shape, bin):
    require_gpu()
    from image_processing_suite_b200 import ops
    F, C, Z, H, W = shape
    _, raw = _fields(F, C, Z, H, W, seed=5)
    raw[0, 0, :, :8, :8] = 65535                       # 16 * 65535 must not overflow uint32
    r = ops.preprocess_fused(dev(raw), None, bin=bin, want_pct_maximal=True)
    mp, binned, pct = host(r["maxproj"]), host(r["binned"]), host(r["pct_maximal"])
    assert binned.dtype == np.uint32
    for f in range(F):
        emp, _, ebin = o_pre.preprocess_field(raw[f], None, bin)
        np.testing.assert_array_equal(mp[f], emp)
        np.testing.assert_array_equal(binned[f], ebin)
        for c in range(C):
            assert pct[f, c] == o_qc.percent_maximal(emp[c])


def test_preprocess_golden_maxproj(golden_dir):
    """The reference's own np.maximum.reduce outputs (tests/golden/maxproj.npz)."""
    require_gpu()
    import os
    from image_processing_suite_b200 import ops
    g = np.load(os.path.join(golden_dir, "maxproj.npz"))
    for k in ("z3", "z5", "z1"):
        planes = g[f"{k}_in"]                           # [Z][H][W]
        r = ops.preprocess_fused(dev(planes[None, None]), None, bin=1, want_binned=False)
        np.testing.assert_array_equal(host(r["maxproj"])[0, 0], g[f"{k}_out"])


def test_preprocess_ragged_width_uses_scalar_path():
    """W % 8 != 0 and odd sizes: the any-shape kernel must give the same answers."""
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 65536, (2, 2, 3, 18, 27), dtype=np.uint16)
    ill = (1.0 + rng.random((2, 18, 27))).astype(np.float32)
    r = ops.preprocess_fused(dev(raw), dev(ill), bin=1, want_corrected=True, want_pct_maximal=True)
    for f in range(2):
        emp, ecorr, ebin = o_pre.preprocess_field(raw[f], ill, 1)
        np.testing.assert_array_equal(host(r["maxproj"])[f], emp)
        np.testing.assert_allclose(host(r["corrected"])[f], ecorr, rtol=RTOL)
    raw3 = rng.integers(0, 65536, (1, 1, 2, 18, 27), dtype=np.uint16)
    r3 = ops.preprocess_fused(dev(raw3), None, bin=1)
    np.testing.assert_array_equal(host(r3["binned"])[0], raw3[0].max(axis=1).astype(np.uint32))


def test_preprocess_errors():
    torch = require_gpu()
    from image_processing_suite_b200 import ops
    raw = torch.zeros((1, 1, 2, 6, 8), dtype=torch.uint16, device="cuda")
    with pytest.raises(ValueError):
        ops.preprocess_fused(raw, None, bin=4)             # 6 % 4 != 0
    with pytest.raises(ValueError):
        ops.preprocess_fused(raw, None, bin=1, want_corrected=True)
    with pytest.raises(ValueError):
        ops.preprocess_fused(raw, torch.ones((1, 6, 9), device="cuda"), bin=1)
    with pytest.raises(ValueError):
        ops.preprocess_fused(raw.cpu(), None)
    assert ops.preprocess_fused(raw[:0], None)["maxproj"].shape[0] == 0   # empty batch


def test_preprocess_full_size_properties():
    """Config-2 size (5ch 2160^2, Z=3): size-independent properties instead of the oracle:
    max >= every plane and equals one of them; sum of bins == sum of max projection;
    bin(4) == bin(2) of bin(2); idempotence (max-projecting the projection)."""
    torch = require_gpu()
    from image_processing_suite_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    raw = torch.randint(0, 65536, (2, 5, 3, 2160, 2160), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
    r2 = ops.preprocess_fused(raw, None, bin=2)
    r4 = ops.preprocess_fused(raw, None, bin=4)
    mp = r2["maxproj"].to(torch.int32)
    rw = raw.to(torch.int32)
    assert bool((mp == rw.amax(dim=2)).all())
    b2 = r2["binned"].view(torch.int32).to(torch.int64)
    b4 = r4["binned"].view(torch.int32).to(torch.int64)
    assert int(b2.sum()) == int(mp.sum(dtype=torch.int64))
    again = b2.reshape(2, 5, 540, 2, 540, 2).sum(dim=(3, 5))
    assert bool((again == b4).all())
    r1 = ops.preprocess_fused(r2["maxproj"][:, :, None].contiguous(), None, bin=1, want_binned=False)
    assert bool((r1["maxproj"].view(torch.int16) == r2["maxproj"].view(torch.int16)).all())


def test_preprocess_tma_variant_passes_the_same_suite():
    """IPS_K1_TMA=1 routes K1 through the TMA-staged kernel (cp.async.bulk + mbarrier ring);
    the flag is read once per process, so the parity cases above are re-run in a child."""
    require_gpu()
    import os
    import subprocess
    import sys
    env = dict(os.environ, IPS_K1_TMA="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_preprocess.py", "-m", "gpu", "-q", "-x",
                        "-k", "not tma_variant"], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
