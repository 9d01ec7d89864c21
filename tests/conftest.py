import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = os.path.join(ROOT, "tests", "golden")

    def load(name):
        return dict(np.load(os.path.join(d, name + ".npz")))

    return load


# The operator-level GPU suites were written against the exact-fp32 CUDA-core engine as their ambient state (each
# test that wants a tensor-core engine selects it and puts "fp32" back).  The LIBRARY default is the tensor-core
# parity engine ("bf16x3", tests/test_cpu_boundary.py::test_default_engine_is_tensor_core_parity); these suites get
# their ambient state set explicitly.  tests/test_gpu_dropin.py runs on the library defaults.
_FP32_AMBIENT = ("test_gpu_parity", "test_gpu_umma", "test_gpu_train", "test_gpu_liif", "test_gpu_ref_sampler")


@pytest.fixture(autouse=True)
def _ambient_engine(request):
    mod = request.module.__name__.rsplit(".", 1)[-1]
    if mod in _FP32_AMBIENT:
        import torch
        if torch.cuda.is_available():
            import anystereo_b200 as A
            A.set_update_engine("fp32")
            A.set_corr_mode("fp32")
    yield
