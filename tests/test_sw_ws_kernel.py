"""The warp-specialised SW two-stream kernel (csrc/kernels/solver_ws.cuh: producer warps compute the two-stream cells
into shared-memory stages, consumer warps run the direct-beam / adding scans) against the CPU oracle and against the
register kernel it shares its arithmetic with: the two must agree BIT FOR BIT on every shape (full tiles, zero-filled
padded tiles, partial last column tile, both orientations, night columns, diffuse boundary condition, broadband and
g-point outputs), and more g-points than pipeline stages so that every stage and barrier phase is reused."""
import numpy as np
import pytest

import refcases as rc
from rte_rrtmgp_b200.abi import fzeros
from test_kernels_parity import _close, _sw_inputs


def _run(lib, device, x, top_at_1, bb, bc):
    ncol, nlay, ngpt = x["tau"].shape
    d = lambda a: rc.dev(a, device)
    if bb:
        decoy = fzeros((ncol, nlay + 1, ngpt), device=device)
        gup = gdn = gdr = decoy
    else:
        gup, gdn, gdr = (fzeros((ncol, nlay + 1, ngpt), device=device) for _ in range(3))
    bup, bdn, bdr = (fzeros((ncol, nlay + 1), device=device) for _ in range(3))
    lib.rte_sw_solver_2stream(ncol, nlay, ngpt, top_at_1, d(x["tau"]), d(x["ssa"]), d(x["g"]), d(x["mu0"]),
                              d(x["adir"]), d(x["adif"]), d(x["inc"]), gup, gdn, gdr, bc, d(x["dif"]), bb, bup, bdn, bdr)
    lib.sync()
    return [rc.host(o) for o in ((bup, bdn, bdr) if bb else (gup, gdn, gdr))]


@pytest.mark.gpu
@pytest.mark.parametrize("top_at_1", [True, False])
@pytest.mark.parametrize("bb,bc", [(True, False), (False, True)])
@pytest.mark.parametrize("ncol,nlay", [(48, 64), (38, 72), (32, 80), (24, 41), (22, 60), (18, 75), (2, 72), (150, 72)])
def test_ws_kernel_matches_register_kernel_and_oracle(oracle_lib, cuda_lib, ncol, nlay, top_at_1, bb, bc):
    x = _sw_inputs(ncol, nlay, 11, seed=100 + ncol + nlay)
    ref = _run(oracle_lib, None, x, top_at_1, bb, bc)
    try:
        cuda_lib.cdll.rrtmgpb_set_solver_variant(2)
        reg = _run(cuda_lib, "cuda:0", x, top_at_1, bb, bc)
        cuda_lib.cdll.rrtmgpb_set_solver_variant(3)
        ws = _run(cuda_lib, "cuda:0", x, top_at_1, bb, bc)
    finally:
        cuda_lib.cdll.rrtmgpb_set_solver_variant(0)
    for a, b, r, n in zip(ws, reg, ref, ("up", "dn", "dir")):
        assert np.all(np.isfinite(a)), n
        _close(a, r, n)
        assert np.array_equal(a, b), f"{n}: warp-specialised and register kernels differ (max {np.max(np.abs(a - b)):.3e})"


@pytest.mark.gpu
def test_ws_kernel_many_gpoints_and_tiles(oracle_lib, cuda_lib):
    """More column tiles than SMs would hold at once is not needed for correctness, but several tiles x 37 g-points
    wrap the four-stage ring nine times and exercise the g-point split of the grid."""
    x = _sw_inputs(200, 72, 37, seed=7)
    ref = _run(oracle_lib, None, x, True, True, False)
    try:
        cuda_lib.cdll.rrtmgpb_set_solver_variant(3)
        ws = _run(cuda_lib, "cuda:0", x, True, True, False)
        ws_g = _run(cuda_lib, "cuda:0", x, True, False, False)
        cuda_lib.cdll.rrtmgpb_set_solver_variant(2)
        reg_g = _run(cuda_lib, "cuda:0", x, True, False, False)
    finally:
        cuda_lib.cdll.rrtmgpb_set_solver_variant(0)
    for a, r, n in zip(ws, ref, ("up", "dn", "dir")):
        _close(a, r, n)
    for a, b in zip(ws_g, reg_g):
        assert np.array_equal(a, b)
