"""smoke(): one small all-sky LW+SW pass on cuda:0 through the product library, checked against the
CPU oracle on identical inputs.  The oracle is imported here only as the checker."""
import numpy as np


def run(ncol=48, nlay=72, verbose=False, tol=1.0e-5):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device")
    import oracle  # checker only
    import rte_rrtmgp_b200 as pkg
    from .allsky import AllSky
    from .frontend import Context
    from . import synthetic as syn

    lib = pkg.lib()
    lib.set_device(0)
    kd_lw, kd_sw = syn.make_kdist("lw"), syn.make_kdist("sw")
    lib.launch_count(reset=True)
    gpu = AllSky(Context(lib, "cuda:0"), ncol, nlay, kd_lw, kd_sw)
    gpu.step()
    fg = gpu.fluxes_host()
    launches = lib.launch_count()
    cpu = AllSky(Context(oracle.lib(), None), ncol, nlay, kd_lw, kd_sw)
    cpu.step()
    fc = cpu.fluxes_host()
    worst = 0.0
    for k in fc:
        err = float(np.max(np.abs(fg[k] - fc[k])))
        worst = max(worst, err)
        if verbose:
            print(f"smoke: {k:12s} max|gpu-oracle| = {err:.3e} W/m2  (max flux {np.max(np.abs(fc[k])):.1f})")
    if verbose:
        print(f"smoke: {launches} kernel launches, worst flux error {worst:.3e} W/m2 (tolerance {tol:g})")
    if not (worst <= tol) or launches <= 0:
        raise AssertionError(f"smoke failed: worst flux error {worst} W/m2, launches {launches}")
    return worst
