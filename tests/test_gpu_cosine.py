"""K4 parity: replicate-group cosine similarity (CUDA) vs scikit-learn-pinned oracle.

cos values live in [-1, 1]: ATOL = 1e-5 on the mean similarity of a group (north_star's
1e-5 against the reference's float64), pair counts exact."""
import os

import numpy as np
import pytest

from oracle import cosine as o_cos
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu
ATOL = 1e-5


@pytest.mark.parametrize("case", ["g4", "g7", "g2", "g1"])
def test_cosine_golden(golden_dir, case):
    require_gpu()
    from image_processing_suite_b200 import ops
    g = np.load(os.path.join(golden_dir, "cosine.npz"))
    x = g[f"{case}_x"]
    s, n = ops.cosine_triu(dev(x.astype(np.float32)))
    ref = g[f"{case}_triu"]
    assert int(host(n)[0]) == ref.size
    if ref.size:
        assert abs(float(host(s)[0]) / ref.size - float(g[f"{case}_mean"])) < ATOL
    else:
        assert float(host(s)[0]) == 0.0


def test_cosine_grouped_matches_oracle():
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(11)
    sizes = [3, 1, 7, 4, 2, 130, 5, 64, 65]
    group = np.repeat(np.arange(len(sizes)), sizes).astype(np.int32)
    x = rng.normal(size=(group.size, 300)).astype(np.float32)
    x[4] = 0.0                                               # a zero row stays zero
    x[10:14] += 5.0                                          # a tight group
    s, n = ops.cosine_triu(dev(x), dev(group))
    ref = o_cos.grouped_triu(x.astype(np.float64), group)
    for gid, (rs, rn) in ref.items():
        assert int(host(n)[gid]) == rn
        if rn:
            assert abs(float(host(s)[gid]) / rn - rs / rn) < ATOL


def test_cosine_single_group_closed_form_large():
    """Size-independent property: sum_{i<j} cos = (|sum x^|^2 - sum |x^|^2) / 2."""
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(12)
    x = rng.normal(size=(3000, 700)).astype(np.float32) + 0.3
    s, n = ops.cosine_triu(dev(x))
    assert int(host(n)[0]) == 3000 * 2999 // 2
    ref = o_cos.triu_sum_closed_form(x.astype(np.float64))
    assert abs(float(host(s)[0]) - ref) / int(host(n)[0]) < ATOL


@pytest.mark.parametrize("n,d", [(1024, 64), (1500, 200), (4100, 700), (2048, 3000)])
def test_cosine_tensor_core_path_matches_closed_form(n, d):
    """One group of >= 1024 rows runs on tcgen05 (bf16 hi/lo split, fp32 TMEM accumulators):
    mean cosine within ATOL of the float64 oracle; exact pair count."""
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(n + d)
    x = (rng.normal(size=(n, d)) + 0.2 * rng.normal(size=(1, d))).astype(np.float32)
    x[7] = 0.0
    s, npairs = ops.cosine_triu(dev(x))
    assert int(host(npairs)[0]) == n * (n - 1) // 2
    ref = o_cos.triu_sum_closed_form(x.astype(np.float64))
    assert abs(float(host(s)[0]) - ref) / (n * (n - 1) // 2) < ATOL
    if n <= 1500:
        assert abs(float(host(s)[0]) - float(o_cos.triu_values(x.astype(np.float64)).sum())) / (n * (n - 1) // 2) < ATOL


def test_cosine_pairs_vector_matches_triu_values():
    """The per-pair vector of every replicate group, in np.triu_indices order."""
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(21)
    sizes = [4, 1, 6, 2, 70, 3]
    group = np.repeat(np.arange(len(sizes)), sizes).astype(np.int32)
    x = rng.normal(size=(group.size, 90)).astype(np.float32)
    s, npairs, pairs, offsets = ops.cosine_triu_pairs(dev(x), dev(group), group_sizes=sizes)
    off = host(offsets)
    assert off.tolist() == np.r_[0, np.cumsum([n * (n - 1) // 2 for n in sizes])].tolist()
    for g, n in enumerate(sizes):
        ref = o_cos.triu_values(x[group == g].astype(np.float64))
        got = host(pairs)[off[g]:off[g + 1]]
        assert got.shape == ref.shape and int(host(npairs)[g]) == ref.size
        np.testing.assert_allclose(got, ref, atol=ATOL)
        if ref.size:
            assert abs(float(host(s)[g]) - ref.sum()) < ATOL * ref.size


def test_cosine_parts_of_the_tile_schedule_add_up():
    """The sharded form on one GPU: the tensor-core pass over every third of the upper-triangular
    tile schedule sums to the whole (ips_cosine_triu_part), and matches the closed form."""
    torch = require_gpu()
    import ctypes as C
    from image_processing_suite_b200 import capi
    from image_processing_suite_b200.cosine_parallel import ShardedCosine
    n, d = 1500, 203
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((n, d), device="cuda", generator=g)
    x[17] = 0                                                   # a zero row stays zero
    sc = ShardedCosine(n, d)
    whole = sc.sum_triu(x)
    parts = []
    part_sum = torch.zeros((1,), dtype=torch.float64, device="cuda")
    for p in range(3):
        capi.call("ips_cosine_triu_part", C.c_void_p(sc.planes.data_ptr()), C.c_void_p(part_sum.data_ptr()), n, d, p, 3, None)
        parts.append(float(part_sum.item()))
    xh = x.double()
    nrm = xh.norm(dim=1, keepdim=True)
    xh = torch.where(nrm > 0, xh / nrm, torch.zeros_like(xh))
    ref = 0.5 * (float((xh.sum(0) ** 2).sum()) - float((xh * xh).sum()))
    npairs = n * (n - 1) / 2
    assert abs(sum(parts) - whole) / npairs < 1e-9
    assert abs(whole - ref) / npairs < 1e-5                     # atol 1e-5 on the mean cosine
    assert all(abs(p) > 0 for p in parts)
    sc.close()
