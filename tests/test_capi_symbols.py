"""The C-ABI library loads and exports every symbol include/ips.h declares (no compute
calls: this runs without a GPU), and the ctypes prototypes cover the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ips.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ips_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from image_processing_suite_b200 import build, capi
    build.build()
    return capi.lib()


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("ips_preprocess_fused", "ips_object_stats", "ips_field_fused", "ips_illum_accumulate",
                 "ips_illum_finalize", "ips_lanczos_resize_u16", "ips_ring_sums", "ips_cosine_triu",
                 "ips_well_mean", "ips_allgather_rows", "ips_pipeline_submit"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, "libips.so lacks: %s" % missing


def test_library_has_no_link_time_nccl_dependency(lib):
    """NCCL is bound with dlopen at run time (csrc/comm.cu): single-GPU users without libnccl
    must still be able to load the library."""
    import subprocess
    from image_processing_suite_b200 import capi
    out = subprocess.run(["readelf", "-d", capi.LIB_PATH], capture_output=True, text=True).stdout
    needed = [l for l in out.splitlines() if "NEEDED" in l]
    assert needed and not any("nccl" in l for l in needed)


def test_prototypes_cover_the_header(lib):
    from image_processing_suite_b200 import capi
    assert sorted(capi.PROTOTYPES) == _declared()
    assert capi.call("ips_abi_version") == 1


def test_entry_points_fail_loudly_without_arguments(lib):
    """Argument validation happens before any CUDA call: errors, not crashes, and no fallback."""
    from image_processing_suite_b200 import capi
    with pytest.raises(capi.IpsError) as e:
        capi.call("ips_preprocess_fused", None, None, None, None, None, 2, None, None, 0, 1, 1, 1, 8, 8, None)
    assert e.value.code == -7 and "raw is NULL" in str(e.value)
    with pytest.raises(capi.IpsError):
        capi.call("ips_object_stats", None, None, None, 1.0, None, None, None, 4, None, 0, 1, 1, 8, 8, None)
    with pytest.raises(capi.IpsError):
        capi.call("ips_lanczos_resize_u16", None, None, 1, 8, 8, 4, 4, None, 0, None)
    assert capi.call("ips_preprocess_workspace_bytes", 1, 5, 2160, 2160, 2) > 0
    assert capi.call("ips_object_stats_workspace_bytes", 16, 5, 2000) >= 16 * 2000 * 20 * 8
    assert capi.call("ips_comm_unique_id_bytes") == 128


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from image_processing_suite_b200 import ops
    with pytest.raises(ValueError, match="no CPU path"):
        ops.preprocess_fused(torch.zeros((1, 1, 1, 8, 8), dtype=torch.uint16))
    from image_processing_suite_b200 import capi
    out = ctypes.c_void_p()
    with pytest.raises(capi.IpsError) as e:
        capi.call("ips_host_alloc", ctypes.byref(out), 1024)
    assert e.value.code == -4                                 # IPS_ERR_CUDA
