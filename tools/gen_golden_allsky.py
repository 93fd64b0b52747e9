#!/usr/bin/env python
"""Write tests/golden/allsky_golden.npz: broadband fluxes of one all-sky LW+SW iteration computed by the CPU
oracle (parity build) for a small seeded case.  The reference itself cannot run here (no Fortran compiler, no
rrtmgp-data), so this fixture does not pin the oracle to the reference - the analytic unit tests do that for
the solvers; it freezes the oracle's answers so that any later change to oracle/, the synthetic generators or
the frontend shows up as a diff, and gives the GPU tests a committed target that does not need the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from rte_rrtmgp_b200 import synthetic as syn  # noqa: E402
from rte_rrtmgp_b200.allsky import AllSky  # noqa: E402
from rte_rrtmgp_b200.frontend import Context  # noqa: E402

CASE = dict(ncol=9, nlay=24, gpt_per_band=2, seed=42)


def compute(lib, device=None):
    kd_lw = syn.make_kdist("lw", gpt_per_band=CASE["gpt_per_band"], seed=CASE["seed"])
    kd_sw = syn.make_kdist("sw", gpt_per_band=CASE["gpt_per_band"], seed=CASE["seed"])
    sky = AllSky(Context(lib, device), CASE["ncol"], CASE["nlay"], kd_lw, kd_sw)
    sky.step()
    return sky.fluxes_host()


if __name__ == "__main__":
    out = compute(oracle.lib())
    path = os.path.join(ROOT, "tests", "golden", "allsky_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path), {k: v.shape for k, v in out.items()})
