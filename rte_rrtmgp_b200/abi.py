"""ctypes access to a shared library that exports the reference's extern-mode kernel ABI.

The same wrapper drives the product library (CUDA, `rte_rrtmgp_b200.lib()`) and, from tests and
the CPU-baseline leg of bench.py only, the CPU oracle.  Calls follow the Fortran convention the
reference's frontend uses: every argument by reference, arrays as base
pointers, Fortran (first-index-fastest) order.

Array arguments may be numpy arrays (host), torch tensors (host or CUDA), or raw integer addresses.
"""
import ctypes

import numpy as np

from ._abi_table import ABI

try:  # torch is plumbing only (device memory); the bindings work without it
    import torch
except Exception:  # pragma: no cover
    torch = None

FLOAT = ctypes.c_double
NP_FLOAT = np.float64

_SCALAR = {"int": ctypes.c_int, "Float": FLOAT, "Bool": ctypes.c_bool}


def _ptr(x):
    """Base address of an array-like argument (numpy / torch / int / None)."""
    if x is None:
        return ctypes.c_void_p(0)
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, np.ndarray):
        return ctypes.c_void_p(x.ctypes.data)
    if torch is not None and isinstance(x, torch.Tensor):
        return ctypes.c_void_p(x.data_ptr())
    if isinstance(x, ctypes.c_void_p):
        return x
    raise TypeError(f"cannot pass {type(x)!r} as an array argument")


class KernelLib:
    """A loaded library exporting the 45 reference kernel symbols (+ the rrtmgpb_* extensions)."""

    def __init__(self, path):
        self.path = path
        # RTLD_LOCAL: the product and the oracle export the SAME symbol names; keep them apart.
        self.cdll = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self._fns = {}
        self.cdll.rrtmgpb_backend_name.restype = ctypes.c_char_p
        self.backend = self.cdll.rrtmgpb_backend_name().decode()
        self.cdll.rrtmgpb_launch_count.restype = ctypes.c_longlong
        self.cdll.rrtmgpb_launch_count.argtypes = [ctypes.c_int]
        self.cdll.rrtmgpb_mem_alloc.restype = ctypes.c_void_p
        self.cdll.rrtmgpb_mem_alloc.argtypes = [ctypes.c_size_t]
        self.cdll.rrtmgpb_mem_free.argtypes = [ctypes.c_void_p]
        self.cdll.rrtmgpb_set_stream.argtypes = [ctypes.c_void_p]
        self.cdll.rrtmgpb_get_stream.restype = ctypes.c_void_p
        self.cdll.rrtmgpb_set_device.argtypes = [ctypes.c_int]
        # working precision of this build (RTE_USE_SP libraries take float32 arrays and c_float scalars)
        self.float_bytes = int(self.cdll.rrtmgpb_float_bytes())
        self.np_float = np.float32 if self.float_bytes == 4 else np.float64
        self._scalar = dict(_SCALAR, Float=ctypes.c_float if self.float_bytes == 4 else ctypes.c_double)

    def symbols(self):
        return sorted(ABI)

    def has(self, name):
        try:
            getattr(self.cdll, name)
            return True
        except AttributeError:
            return False

    def call(self, name, *args):
        """Call reference-ABI symbol `name` with positional args in the Fortran argument order."""
        sig = ABI[name]
        if len(args) != len(sig):
            raise TypeError(f"{name} takes {len(sig)} arguments ({len(args)} given)")
        fn = self._fns.get(name)
        if fn is None:
            fn = getattr(self.cdll, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_void_p] * len(sig)
            self._fns[name] = fn
        keep, cargs = [], []
        for (aname, ctype, is_arr, _intent), val in zip(sig, args):
            if is_arr:
                cargs.append(_ptr(val))
            else:
                box = self._scalar[ctype](val)
                keep.append(box)
                cargs.append(ctypes.cast(ctypes.pointer(box), ctypes.c_void_p))
        fn(*cargs)

    def __getattr__(self, name):
        if name in ABI:
            return lambda *a: self.call(name, *a)
        raise AttributeError(name)

    # ---- plumbing ----
    def cast(self, a):
        """Floating-point numpy arrays in this build's working precision (anything else unchanged)."""
        if isinstance(a, np.ndarray) and a.dtype.kind == "f" and a.dtype != self.np_float:
            return np.asfortranarray(a, dtype=self.np_float)
        return a

    def sync(self):
        self.cdll.rrtmgpb_sync()

    def launch_count(self, reset=False):
        return int(self.cdll.rrtmgpb_launch_count(1 if reset else 0))

    def set_stream(self, stream_handle):
        self.cdll.rrtmgpb_set_stream(ctypes.c_void_p(stream_handle))

    def set_device(self, dev):
        self.cdll.rrtmgpb_set_device(int(dev))


# ---- Fortran-ordered array helpers -------------------------------------------------------------
def fzeros(shape, dtype=NP_FLOAT, device=None):
    """Zero array with Fortran shape `shape`, first index fastest.  numpy if device is None,
    else a torch tensor on `device` (a permuted view of a C-contiguous tensor)."""
    if device is None:
        return np.zeros(shape, dtype=dtype, order="F")
    tdt = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32, np.bool_: torch.bool}[
        np.dtype(dtype).type
    ]
    t = torch.zeros(tuple(reversed(shape)), dtype=tdt, device=device)
    return t.permute(*reversed(range(len(shape))))


def to_device(a, device):
    """Copy a numpy F-ordered array to `device`, keeping the Fortran memory order."""
    a = np.asfortranarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(device)
    return t.permute(*reversed(range(a.ndim)))


def to_host(t):
    """Inverse of to_device: torch F-view (or numpy) -> numpy F-ordered array."""
    if isinstance(t, np.ndarray):
        return t
    nd = t.dim()
    base = t.permute(*reversed(range(nd))).contiguous().cpu().numpy()
    return np.asfortranarray(base.T)
