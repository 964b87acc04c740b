import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_oracle_port():
    so = os.path.join(ROOT, "oracle", "libppx_oracle.so")
    src = os.path.join(ROOT, "oracle", "ppx_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def port():
    """The C restatement (oracle/ppx_oracle.c), strict IEEE build."""
    from oracle_lib import Oracle
    _ensure_oracle_port()
    return Oracle("port")


@pytest.fixture(scope="session")
def ref_strict():
    """The reference headers behind ref_shim.cpp (-ffp-contract=off); skipped when not built."""
    from oracle_lib import Oracle, available
    if not available("ref_strict"):
        pytest.skip("oracle/_ref/libppx_ref_strict.so not built (needs /root/reference)")
    return Oracle("ref_strict")


@pytest.fixture(scope="session", params=["port", "ref_strict"])
def cpu_ref(request):
    """The checker of the headline parity tests, once as the C restatement and once as the compiled reference itself
    (oracle/_ref/libppx_ref_strict.so: the unmodified ppx headers behind ref_shim.cpp; it ships prebuilt to the GPU
    box).  The two are pinned bit-for-bit to each other by tests/test_oracle.py."""
    from oracle_lib import Oracle, available
    if request.param == "port":
        _ensure_oracle_port()
        return Oracle("port")
    if not available("ref_strict"):
        pytest.skip("oracle/_ref/libppx_ref_strict.so not built (needs /root/reference)")
    return Oracle("ref_strict")


@pytest.fixture(scope="session")
def ref():
    from oracle_lib import Oracle, available
    if not available("ref"):
        pytest.skip("oracle/_ref/libppx_ref.so not built (needs /root/reference)")
    return Oracle("ref")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixtures():
    return dict(np.load(os.path.join(HERE, "golden", "ref_outputs.npz")))
