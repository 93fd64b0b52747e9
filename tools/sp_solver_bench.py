#!/usr/bin/env python
"""Single-precision build (lib/librte_rrtmgp_b200_sp.so, the reference's RTE_ENABLE_SP): device times of the two flux
solvers through the extern ABI at the headline shape (65,536 x 72, 256 / 224 g-points, broadband outputs), next to the
same call on the double-precision library.  Prints one JSON line; used by bench.py as a supplementary key."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(lib, ncol, nlay, steps):
    import torch

    from rte_rrtmgp_b200.abi import fzeros

    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    tdt = torch.float32 if lib.float_bytes == 4 else torch.float64
    F = lib.np_float
    gen = torch.Generator(device="cuda").manual_seed(1)

    def plane(*shape, lo=0.0, hi=1.0):   # Fortran (first-index-fastest) view of a C-contiguous tensor
        t = torch.rand(tuple(reversed(shape)), dtype=tdt, device="cuda", generator=gen) * (hi - lo) + lo
        return t.permute(*reversed(range(len(shape))))

    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for name, ngpt in (("lw_noscat", 256), ("sw_2stream", 224)):
        tau = plane(ncol, nlay, ngpt, lo=0.001, hi=2.0)
        decoy = fzeros((1,), F, "cuda:0")
        bup, bdn, bdr = (fzeros((ncol, nlay + 1), F, "cuda:0") for _ in range(3))
        if name == "lw_noscat":
            lay, lev = plane(ncol, nlay, ngpt, lo=50, hi=100), plane(ncol, nlay + 1, ngpt, lo=50, hi=100)
            emis, sfc, inc = plane(ncol, ngpt, lo=0.9, hi=1.0), plane(ncol, ngpt, lo=80, hi=100), plane(ncol, ngpt)
            Ds, w = plane(ncol, ngpt, lo=1.6, hi=1.7), torch.ones(1, dtype=tdt, device="cuda")
            call = lambda: lib.rte_lw_solver_noscat(ncol, nlay, ngpt, True, 1, Ds, w, tau, lay, lev, emis, sfc, inc, decoy, decoy,
                                                    True, bup, bdn, False, decoy, decoy, False, decoy, decoy)
            nbytes = (2 * ncol * nlay + ncol * (nlay + 1)) * ngpt * lib.float_bytes
        else:
            ssa, g = plane(ncol, nlay, ngpt, lo=0.1, hi=0.99), plane(ncol, nlay, ngpt, lo=0.0, hi=0.9)
            mu0, ad, af, inc = plane(ncol, nlay, lo=0.3, hi=1.0), plane(ncol, ngpt, hi=0.5), plane(ncol, ngpt, hi=0.5), plane(ncol, ngpt, hi=5.0)
            call = lambda: lib.rte_sw_solver_2stream(ncol, nlay, ngpt, True, tau, ssa, g, mu0, ad, af, inc, decoy, decoy, decoy, False,
                                                     decoy, True, bup, bdn, bdr)
            nbytes = 3 * ncol * nlay * ngpt * lib.float_bytes
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(steps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms": round(ms, 3), "input_GB": round(nbytes / 1e9, 2), "GBps": round(nbytes / ms / 1e6, 1)}
        del tau
        torch.cuda.empty_cache()
    return out


def main():
    import rte_rrtmgp_b200 as pkg

    ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    res = {"ncol": ncol, "nlay": 72, "sp": run(pkg.lib_sp(), ncol, 72, 5), "dp": run(pkg.lib(), ncol, 72, 5),
           "note": "rte_lw_solver_noscat / rte_sw_solver_2stream (broadband) through the extern ABI, device arrays, random planes"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
