"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/qilqr.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qilqr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qilqr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from quadrotorilqr_b200 import _capi

    _capi.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_capi.EXPORTED_SYMBOLS) == names


def test_library_targets_sm_100a_only():
    from quadrotorilqr_b200 import _capi

    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_error_strings_and_structs():
    from quadrotorilqr_b200 import _capi

    lib = _capi.lib()
    assert lib.qilqr_error_string(0) == b"ok"
    assert b"positive definite" in lib.qilqr_error_string(_capi.ERR_INERTIA_NOT_PD)
    assert ctypes.sizeof(_capi.Result) == 24
    assert ctypes.sizeof(_capi.Options) == 64


def test_create_rejects_bad_inertia_before_touching_the_gpu():
    """quadrotor_model.cc:21-24 -> QILQR_ERR_INERTIA_NOT_PD (host-side check, no device needed)."""
    import numpy as np
    from quadrotorilqr_b200 import BatchILQR, QilqrError, _capi

    with pytest.raises(QilqrError) as e:
        BatchILQR(1.0, np.diag([1.0, -1.0, 1.0]), 1.0, 0.0, 9.81, np.eye(12), np.eye(4), 0.1)
    assert e.value.code == _capi.ERR_INERTIA_NOT_PD


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly, not compute on the CPU."""
    import numpy as np
    import torch
    from quadrotorilqr_b200 import BatchILQR, QilqrError, _capi

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(QilqrError) as e:
        BatchILQR(1.0, np.eye(3), 1.0, 0.0, 9.81, np.eye(12), np.eye(4), 0.1)
    assert e.value.code == _capi.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "quadrotorilqr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "qoracle" not in text, f
