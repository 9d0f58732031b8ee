"""CPU oracle for the per-image hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a NumPy / SciPy / Pillow restatement of the arithmetic that
Saguaro-Biosciences/image-processing-suite performs on its per-image hot path
(SURVEY.md section 8a rows A1-A8).  It exists so that the CUDA path can be checked
against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; nothing under
``image_processing_suite_b200/`` does, and the product path raises if the CUDA
extension is missing rather than falling back to anything in here.

Pinning status (see DESIGN.md "Oracle"):
  * A1 max projection, A2 illumination divide, A6 PercentMaximal, A7 rps /
    PowerLogLogSlope, A3 Pillow-LANCZOS: PINNED -- ``oracle/make_golden.py`` imports
    the reference's own functions from /root/reference (with boto3/imageio/tifffile
    stubbed) and the real Pillow, and the restatements here are checked against those
    outputs (``tests/golden/*.npz``, ``tests/test_oracle_golden.py``).
  * A8 cosine: pinned against scikit-learn's ``cosine_similarity`` (the function the
    reference calls), and with the double sigmoid against the files the reference's own
    ``Feature_select_cosine_ami`` loop writes (``tests/golden/cosine_script.npz``).
  * well aggregation (``normalize.well_mean``): PINNED against the per-well table the
    reference's own ``Normalize_CP_ami.concatenate_csv_from_s3`` writes when pycytominer's
    ``normalize`` is stubbed to the identity (``tests/golden/well_agg.npz``).
  * A3' sum binning, A4 per-object statistics, A5 illumination estimation,
    mad_robustize: PARITY UNPINNED -- the reference holds no such arithmetic (it is
    north_star-only, or lives in CellProfiler 4.2.8 / pycytominer, neither vendored).
    The restatements follow SURVEY.md section 8c.
"""
from . import preprocess, object_stats, illum, lanczos, qc, cosine, normalize, crops  # noqa: F401
