"""Host-buffer pipeline (ips_pipeline_*): pinned host in, pinned host out, vs the oracle."""
import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import object_stats as o_obj
from oracle import preprocess as o_pre
from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("with_illum,label_dtype,W", [(True, np.int32, 128), (False, np.int32, 128), (True, np.uint16, 128),
                                                      (False, np.uint16, 128), (True, np.uint16, 100)])
def test_pipeline_batches_match_oracle(with_illum, label_dtype, W):
    """W = 128: the packed-label kernel (uint16 masks as they are, reciprocal function); W = 100:
    the general kernels behind the same call (masks widened on the device)."""
    require_gpu()
    from image_processing_suite_b200.pipeline import FieldPipeline, pinned_empty
    Fb, C, Z, H, cells, n_batches = 2, 3, 3, 96, 12, 5
    ill = synth.make_illum(C, H, W, seed=5) if with_illum else None
    scale = 1.0 / 65535.0 if with_illum else 1.0
    pipe = FieldPipeline(Fb, C, Z, H, W, bin=2, n_max=cells, depth=2, illum=ill, intensity_scale=scale,
                         label_dtype=label_dtype)
    raws, labs, outs, tickets = [], [], [], []
    for b in range(n_batches):
        raw = pinned_empty((Fb, C, Z, H, W), np.uint16)
        lab = pinned_empty((Fb, H, W), label_dtype)
        for k in range(Fb):
            lab[k] = synth.make_labels(H, W, cells, seed=10 * b + k, amin=5, amax=10)
            raw[k] = synth.field_numpy(lab[k], c=C, z=Z, seed=10 * b + k)
        out = pipe.output_buffers()
        raws.append(raw); labs.append(lab); outs.append(out)
        tickets.append(pipe.submit(raw, lab, out))          # more batches in flight than slots
    for t in tickets:
        pipe.wait(t)
    assert tickets == list(range(n_batches))
    for b in range(n_batches):
        for k in range(Fb):
            mp, _, binned = o_pre.preprocess_field(raws[b][k], ill, 2)
            np.testing.assert_array_equal(outs[b]["maxproj"][k], mp)
            if with_illum:
                np.testing.assert_allclose(outs[b]["binned"][k], binned, rtol=1e-5)
            else:
                np.testing.assert_array_equal(outs[b]["binned"][k], binned)
            e_i, e_f = o_obj.object_stats(labs[b][k], mp, ill, scale)
            n = int(outs[b]["n_objects"][k])
            assert n == e_i.shape[0]
            np.testing.assert_array_equal(outs[b]["ints"][k, :n], e_i)
            np.testing.assert_allclose(outs[b]["flts"][k, :n], e_f, rtol=1e-5, atol=1e-6)
    pipe.close()


def test_pipeline_argument_errors():
    require_gpu()
    from image_processing_suite_b200 import capi
    from image_processing_suite_b200.pipeline import FieldPipeline
    with pytest.raises(capi.IpsError):
        FieldPipeline(1, 3, 3, 30, 32, bin=4)               # 30 % 4 != 0
    pipe = FieldPipeline(1, 1, 1, 16, 16, bin=1, n_max=4)
    with pytest.raises(ValueError):
        pipe.submit(np.zeros((1, 1, 1, 16, 8), np.uint16), np.zeros((1, 16, 16), np.int32), {})
    with pytest.raises(capi.IpsError):
        pipe.wait(7)
    pipe.close()
