"""rte_rrtmgp_b200 - Blackwell-native (sm_100a) RTE+RRTMGP compute path.

Product layout:
  csrc/kernels/   hand-written CUDA kernels
  csrc/abi/       extern "C" entry points = the reference's RTE_KERNEL_MODE=extern symbols
  csrc/frontend/  C++ mirror of the reference's Fortran frontend for this path
  lib/            the built shared library (git-ignored, travels to the GPU box)
This Python package is plumbing: ctypes bindings, torch for device memory / streams / distributed.
"""
import os

from .abi import KernelLib, fzeros, to_device, to_host  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librte_rrtmgp_b200.so")
_LIB = None


def lib():
    """The product library.  Fails loudly if the CUDA extension has not been built - there is no
    CPU fallback."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("RRTMGPB_LIB", LIB_PATH)  # A/B measurements of build variants (tools/build_variant.py)
        if path != LIB_PATH:
            _LIB = KernelLib(path)
            return _LIB
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  rte_rrtmgp_b200 has no CPU fallback."
            )
        _LIB = KernelLib(LIB_PATH)
        if not _LIB.backend.startswith("cuda"):
            raise RuntimeError(f"{LIB_PATH} reports backend {_LIB.backend!r}; expected the CUDA build")
    return _LIB
