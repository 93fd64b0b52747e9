#!/usr/bin/env python
"""What a stock host gets (a gfortran build of the reference linked against this library, arrays in HOST memory): the
all-sky step driven through the 45 extern symbols one by one with numpy arrays - every argument of every call is staged
to the device and back (csrc/common.cuh DevArg).  Correctness path, PCIe-bound; prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np

    import oracle
    import rte_rrtmgp_b200 as pkg
    from rte_rrtmgp_b200 import synthetic as syn
    from rte_rrtmgp_b200.allsky import AllSky
    from rte_rrtmgp_b200.frontend import Context

    lib = pkg.lib()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    kd_lw, kd_sw = syn.make_kdist("lw"), syn.make_kdist("sw")
    sky = AllSky(Context(lib, None), n, 72, kd_lw, kd_sw, fused=False)   # device=None: numpy arrays into the CUDA library
    sky.step()
    lib.sync()
    t0 = time.perf_counter()
    sky.step()
    lib.sync()
    dt = time.perf_counter() - t0
    chk = AllSky(Context(oracle.lib(), None), 48, 72, kd_lw, kd_sw)
    chk.step()
    fg, fc = sky.fluxes_host(), chk.fluxes_host()
    err = max(float(np.max(np.abs(fg[k][:48] - fc[k]))) for k in fc)
    print(json.dumps({"value": n / dt, "unit": "columns/s", "columns": n, "seconds": dt, "max_abs_flux_err_vs_oracle_Wm2": err,
                      "note": f"{n} columns, numpy (pageable host) arrays handed to the 45 symbols one by one: every argument is "
                              f"staged to the device and back per call; max flux error vs oracle {err:.2e} W/m2"}))


if __name__ == "__main__":
    main()
