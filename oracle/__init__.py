"""ORACLE - TEST INFRASTRUCTURE ONLY.

Python loader for the CPU restatement of the reference kernels (oracle/*.c).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Nothing under rte_rrtmgp_b200/ imports it (tests/test_abi.py checks).
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def _name(fast, sp):
    return "liboracle_sp.so" if sp else ("liboracle_fast.so" if fast else "liboracle.so")


def build(fast=False, sp=False):
    target = "sp" if sp else ("fast" if fast else "all")
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)
    return os.path.join(_HERE, "_build", _name(fast, sp))


def lib(fast=False, sp=False):
    """KernelLib over the oracle (parity build by default; `fast` = -O3 -march=native build; `sp` = the parity build in
    single precision, -DRTE_USE_SP)."""
    key = (bool(fast), bool(sp))
    if key not in _LIBS:
        from rte_rrtmgp_b200.abi import KernelLib

        path = os.path.join(_HERE, "_build", _name(fast, sp))
        srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
        front = os.path.join(_HERE, "..", "rte_rrtmgp_b200", "csrc", "frontend")
        if os.path.isdir(front):
            srcs += [os.path.join(front, f) for f in os.listdir(front)]
        stale = (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs)
        if stale:
            build(fast, sp)
        _LIBS[key] = KernelLib(path)
    return _LIBS[key]
