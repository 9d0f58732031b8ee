"""K5 parity: Pillow-exact LANCZOS resize (CUDA) vs Pillow itself and the golden fixtures
produced by the reference's process_image_in_memory arithmetic.  Bit-exact."""
import os

import numpy as np
import pytest

from oracle import lanczos as o_lz
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["x2", "x4", "odd", "full"])
def test_lanczos_golden(golden_dir, case):
    require_gpu()
    from image_processing_suite_b200 import ops
    g = np.load(os.path.join(golden_dir, "rebin_lanczos.npz"))
    src, ref = g[f"{case}_in"], g[f"{case}_out"]
    got = host(ops.lanczos_resize_u16(dev(src[None]), ref.shape))[0]
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("shape,out", [((2, 216, 216), (108, 108)), ((1, 216, 216), (54, 54)), ((2, 216, 432), (72, 144)), ((1, 215, 431), (215, 431 // 1)),
                                       ((2, 300, 201), (150, 67)), ((1, 256, 256), (64, 128)),
                                       ((3, 100, 150), (37, 61)), ((1, 64, 64), (64, 32)),
                                       ((1, 64, 64), (32, 64)), ((1, 40, 40), (40, 40)),
                                       ((1, 30, 30), (75, 45))])
def test_lanczos_matches_pillow(shape, out):
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(7)
    src = rng.integers(0, 65536, shape, dtype=np.uint16)
    src[0, :10, :10] = 65535                              # saturated plateau: overflow quirk
    got = host(ops.lanczos_resize_u16(dev(src), out))
    for c in range(shape[0]):
        np.testing.assert_array_equal(got[c], o_lz.pil_resize(src[c], out))


def test_lanczos_full_field_2160_to_1080_and_540():
    """The script's defaults: 2160^2 -> --resolution 1080 and 540, against Pillow."""
    require_gpu()
    from image_processing_suite_b200 import ops
    rng = np.random.default_rng(8)
    src = rng.integers(0, 4096, (1, 2160, 2160), dtype=np.uint16)
    src[0, 100:140, 200:260] = 65535
    d = dev(src)
    for res in (1080, 540):
        np.testing.assert_array_equal(host(ops.lanczos_resize_u16(d, (res, res)))[0],
                                      o_lz.pil_resize(src[0], (res, res)))
