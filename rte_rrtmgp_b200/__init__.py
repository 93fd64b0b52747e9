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
LIB_PATH_SP = os.path.join(_HERE, "lib", "librte_rrtmgp_b200_sp.so")
_LIB = None
_LIB_SP = None


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


def lib_sp():
    """The single-precision build of the product library (-DRTE_USE_SP, the reference's RTE_ENABLE_SP): the same 45
    symbols on float32 arrays.  No CPU fallback either."""
    global _LIB_SP
    if _LIB_SP is None:
        if not os.path.exists(LIB_PATH_SP):
            raise RuntimeError(f"{LIB_PATH_SP} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB_SP = KernelLib(LIB_PATH_SP)
        if not _LIB_SP.backend.startswith("cuda") or _LIB_SP.float_bytes != 4:
            raise RuntimeError(f"{LIB_PATH_SP}: expected the single-precision CUDA build")
    return _LIB_SP
