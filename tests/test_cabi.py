"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/multibox_b200.h declares (no compute calls here:
this container has no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "multibox_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbx_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("mbx_version", "mbx_last_error", "mbx_match_loss", "mbx_match_workspace_bytes",
              "mbx_detect", "mbx_detect_workspace_bytes", "mbx_filter_proposals", "mbx_convert_proposals"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    from multibox_b200 import _build, _lib
    path = _build.build()
    assert os.path.isfile(path)
    lib = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(lib, s), "missing export %s" % s
    assert set(_declared_symbols()) == set(_lib.EXPORTS)
    lib.mbx_version.restype = ctypes.c_int
    assert lib.mbx_version() == 100
    lib.mbx_match_workspace_bytes.restype = ctypes.c_size_t
    assert lib.mbx_match_workspace_bytes(32, 646, 20) >= 32 * 20


def test_sass_is_sm100a_with_tma_bulk_copy():
    import shutil
    import subprocess
    from multibox_b200 import _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _build.build()], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out          # cp.async.bulk (TMA) staging of the priors


def test_no_cpu_fallback_in_product():
    # the product never imports the oracle and has no numpy/scipy solver path
    pkg = os.path.join(ROOT, "multibox_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/", "").replace("the CPU oracle", ""), fn
            assert "linear_sum_assignment(" not in src, fn


def test_cuda_entry_points_refuse_cpu_tensors():
    torch = pytest.importorskip("torch")
    from multibox_b200 import loss
    with pytest.raises(TypeError):
        loss.compute_assignments(torch.zeros(4, 4), torch.zeros(4), torch.zeros(1, 1, 4),
                                 torch.zeros(1, dtype=torch.int32), 1, 1.0)
