import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_lib():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import rte_rrtmgp_b200

    return rte_rrtmgp_b200.lib()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    return oracle.lib()


@pytest.fixture(scope="session")
def cuda_lib():
    return _cuda_lib()


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu),
                        pytest.param("cuda-tile", marks=pytest.mark.gpu)])
def backend(request):
    """(KernelLib, device) for the CPU oracle and - under `-m gpu` - for the CUDA product library, once with
    the default (register / warp-systolic) solver kernels and once forced onto the shared-memory tile kernels."""
    if request.param == "oracle":
        import oracle

        yield oracle.lib(), None
        return
    lib = _cuda_lib()
    lib.cdll.rrtmgpb_set_solver_variant(1 if request.param == "cuda-tile" else 0)
    yield lib, "cuda:0"
    lib.cdll.rrtmgpb_set_solver_variant(0)
