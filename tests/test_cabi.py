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
    """The product never imports / loads / executes the oracle and has no numpy/scipy solver path:
    checked on the syntax tree (imports, string literals, calls), not by grepping words."""
    import ast
    pkg = os.path.join(ROOT, "multibox_b200")
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read(), fn)
        docstrings = set()
        for node in ast.walk(tree):
            if isinstance(node, (ast.Module, ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)) and node.body and \
                    isinstance(node.body[0], ast.Expr) and isinstance(node.body[0].value, ast.Constant):
                docstrings.add(id(node.body[0].value))
        for node in ast.walk(tree):
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            else:
                mods = []
            for m in mods:
                root = m.split(".")[0]
                assert root not in ("oracle", "scipy", "sklearn", "numba"), (fn, m)
            if isinstance(node, ast.Constant) and isinstance(node.value, str) and id(node) not in docstrings:
                # a path into oracle/ in a string that is not a docstring would be a load / exec of the checker
                assert "oracle/" not in node.value and "libmbx_oracle" not in node.value, (fn, node.value[:60])
            if isinstance(node, ast.Call):
                name = getattr(node.func, "attr", getattr(node.func, "id", ""))
                assert name not in ("linear_sum_assignment", "exec", "eval"), (fn, name)
                if name == "CDLL":      # the only library the product loads is its own
                    assert fn == "_lib.py", fn
    # subprocesses: only the nvcc build
    for fn in sorted(os.listdir(pkg)):
        if fn.endswith(".py") and "subprocess" in open(os.path.join(pkg, fn)).read():
            assert fn == "_build.py", fn


def test_native_boundary_has_no_cpu_path():
    """The py_func drop-in (host numpy in / out) must fail loudly without a CUDA device."""
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_boundary.py")
    import numpy as np
    from multibox_b200 import native_boundary
    with pytest.raises(RuntimeError, match="no CPU path"):
        native_boundary.compute_assignments(np.zeros((4, 4), np.float32), np.full(4, 0.5, np.float32),
                                            np.zeros((1, 1, 4), np.float32), np.ones(1, np.int32), 1, 1.0)


def test_cuda_entry_points_refuse_cpu_tensors():
    torch = pytest.importorskip("torch")
    from multibox_b200 import loss
    with pytest.raises(TypeError):
        loss.compute_assignments(torch.zeros(4, 4), torch.zeros(4), torch.zeros(1, 1, 4),
                                 torch.zeros(1, dtype=torch.int32), 1, 1.0)
