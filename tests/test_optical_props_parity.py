"""GPU parity of the optical-properties arithmetic (rte/kernels/mo_optical_props_kernels.F90:47-706; api
rte/kernels/api/mo_optical_props_kernels.F90:36-368): all 18 increments, both delta scalings and the three extractors,
CUDA through the C-ABI vs the oracle on seeded random multi-band data of shape (37 columns, 19 layers, 3 bands / 11
g-points), with nmom1 != nmom2 for the n-stream pairs and optically empty cells (tau = 0) so that the eps = 3*tiny
guards (:38) are exercised.  Tolerance 1e-13 relative (division and FMA contraction differ in the last bits), plus 1e-14
absolute on g and the moments, whose numerators can cancel."""
import numpy as np
import pytest

import refcases as rc
from rte_rrtmgp_b200.abi import fzeros

NCOL, NLAY, NGPT, NBND = 37, 19, 11, 3
LIMS = np.asfortranarray(np.array([[1, 5, 6], [4, 5, 11]], dtype=np.int32))  # bands of 4, 1 and 6 g-points
RTOL = 1.0e-13
KINDS = ["1scalar", "2stream", "nstream"]


def _props(rng, kind, ngpt, nmom):
    tau = 10.0 ** rng.uniform(-8, 1.5, (NCOL, NLAY, ngpt))
    tau[rng.random(tau.shape) < 0.05] = 0.0
    ssa = rng.uniform(0.0, 1.0, (NCOL, NLAY, ngpt))
    ssa[rng.random(ssa.shape) < 0.05] = 0.0
    out = [np.asfortranarray(tau)]
    if kind != "1scalar":
        out.append(np.asfortranarray(ssa))
    if kind == "2stream":
        out.append(np.asfortranarray(rng.uniform(-0.5, 0.95, (NCOL, NLAY, ngpt))))
    if kind == "nstream":
        out.append(np.asfortranarray(rng.uniform(-0.5, 0.95, (nmom, NCOL, NLAY, ngpt))))
    return out


def _call(lib, device, k1, k2, bybnd, nmom1, nmom2, op1, op2):
    d = lambda a: rc.dev(a.copy(order="F"), device)
    a1, a2 = [d(a) for a in op1], [d(a) for a in op2]
    u1, u2 = list(a1), list(a2)
    if k1 == "1scalar" and k2 != "1scalar":
        u2 = u2[:2]          # tau1 += tau2*(1 - ssa2)                                   :137-139, :158-160
    if k1 != "1scalar" and k2 == "1scalar":
        u1 = u1[:2]          # g / p unchanged                                           :181-190
    sizes = [NCOL, NLAY, NGPT]
    if k1 == "nstream" and k2 != "1scalar":
        sizes.append(nmom1)
    if k2 == "nstream" and k1 != "1scalar":
        sizes.append(nmom2)
    if bybnd:
        lib.call(f"rte_inc_{k1}_by_{k2}_bybnd", *sizes, *u1, *u2, NBND, rc.dev(LIMS, device))
    else:
        lib.call(f"rte_increment_{k1}_by_{k2}", *sizes, *u1, *u2)
    lib.sync()
    return [rc.host(a) for a in a1]


@pytest.mark.gpu
@pytest.mark.parametrize("bybnd", [False, True])
@pytest.mark.parametrize("k2", KINDS)
@pytest.mark.parametrize("k1", KINDS)
@pytest.mark.parametrize("nmom1,nmom2", [(4, 2), (2, 5), (3, 3)])
def test_increment_vs_oracle(oracle_lib, cuda_lib, k1, k2, bybnd, nmom1, nmom2):
    if "nstream" not in (k1, k2) and (nmom1, nmom2) != (3, 3):
        pytest.skip("moment counts only matter for n-stream operands")
    rng = np.random.default_rng(100 * KINDS.index(k1) + 10 * KINDS.index(k2) + bybnd)
    op1 = _props(rng, k1, NGPT, nmom1)
    op2 = _props(rng, k2, NBND if bybnd else NGPT, nmom2)
    ref = _call(oracle_lib, None, k1, k2, bybnd, nmom1, nmom2, op1, op2)
    got = _call(cuda_lib, "cuda:0", k1, k2, bybnd, nmom1, nmom2, op1, op2)
    for a, b, o, n in zip(got, ref, op1, ("tau", "ssa", "g/p")):
        # tau and ssa are sums / quotients of non-negative terms: purely relative.  g and the phase-function moments
        # combine terms of either sign (:213-222, :274-300): where they cancel, the last-bit differences of the two
        # products are an ABSOLUTE error of a few 1e-16 on an O(1) quantity
        atol = 1e-300 if n != "g/p" else 1.0e-14
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=atol, err_msg=f"{k1}+={k2} bybnd={bybnd}: {n}")
    assert not np.array_equal(ref[0], op1[0])  # the increment did something


@pytest.mark.gpu
def test_delta_scale_vs_oracle(oracle_lib, cuda_lib):
    rng = np.random.default_rng(7)
    tau, ssa, g = _props(rng, "2stream", NGPT, 0)
    f = np.asfortranarray(rng.uniform(0.0, 0.9, tau.shape))
    f[rng.random(f.shape) < 0.05] = 1.0   # (1 - f) -> 0: max(eps, .) guard of :69
    res = {}
    for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
        d = lambda a: rc.dev(a.copy(order="F"), device)
        t1, s1, g1 = d(tau), d(ssa), d(g)
        lib.rte_delta_scale_2str_k(NCOL, NLAY, NGPT, t1, s1, g1)
        t2, s2, g2 = d(tau), d(ssa), d(g)
        lib.rte_delta_scale_2str_f_k(NCOL, NLAY, NGPT, t2, s2, g2, d(f))
        lib.sync()
        res[name] = [rc.host(a) for a in (t1, s1, g1, t2, s2, g2)]
    for a, b, n in zip(res["gpu"], res["ref"], ("tau", "ssa", "g", "tau_f", "ssa_f", "g_f")):
        # (g - f)/(1 - f), (ssa - ssa*f)/(1 - ssa*f): differences -> absolute tolerance on O(1) quantities
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-300 if n.startswith("tau") else 1.0e-14, err_msg=n)


@pytest.mark.gpu
def test_extract_subsets_vs_oracle(oracle_lib, cuda_lib):
    rng = np.random.default_rng(8)
    tau, ssa = _props(rng, "2stream", NGPT, 0)[:2]
    p = _props(rng, "nstream", NGPT, 4)[2]
    colS, colE = 6, 29
    n = colE - colS + 1
    res = {}
    for name, lib, device in (("ref", oracle_lib, None), ("gpu", cuda_lib, "cuda:0")):
        d = lambda a: rc.dev(a, device)
        o1 = fzeros((n, NLAY, NGPT), device=device)
        lib.rte_extract_subset_dim1_3d(NCOL, NLAY, NGPT, d(tau), colS, colE, o1)
        o2 = fzeros((4, n, NLAY, NGPT), device=device)
        lib.rte_extract_subset_dim2_4d(4, NCOL, NLAY, NGPT, d(p), colS, colE, o2)
        o3 = fzeros((n, NLAY, NGPT), device=device)
        lib.rte_extract_subset_absorption_tau(NCOL, NLAY, NGPT, d(tau), d(ssa), colS, colE, o3)
        lib.sync()
        res[name] = [rc.host(a) for a in (o1, o2, o3)]
    assert np.array_equal(res["gpu"][0], res["ref"][0]) and np.array_equal(res["ref"][0], tau[colS - 1:colE])
    assert np.array_equal(res["gpu"][1], res["ref"][1]) and np.array_equal(res["ref"][1], p[:, colS - 1:colE])
    np.testing.assert_allclose(res["gpu"][2], res["ref"][2], rtol=RTOL, atol=1e-300)
