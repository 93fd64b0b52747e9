"""Restatement of the reference's tests/rte_optic_prop_unit_tests.F90 at kernel level (:72-190):
incrementing by a transparent medium leaves properties unchanged (all 9 type pairs), two media of
half the optical thickness add up to the original, delta-scaling with f = 0 is the identity.
`ops_match` = allclose with 2 spacings (tests/mo_comparisons.F90:164-218)."""
import numpy as np
import pytest

import refcases as rc
from rte_rrtmgp_b200.abi import fzeros

NCOL, NLAY, NMOM, NGPT = 4, 8, 4, 1
TOTAL_TAU = np.array([0.1, 1.0, 10.0, 50.0])
G, SSA = 0.85, 1.0 - 1.0e-4


def _ref():
    tau = np.zeros((NCOL, NLAY, NGPT), order="F")
    tau[:, :, 0] = (TOTAL_TAU / NLAY)[:, None]
    ssa = np.full_like(tau, SSA)
    g = np.full_like(tau, G)
    p = np.zeros((NMOM, NCOL, NLAY, NGPT), order="F")
    for m in range(NMOM):
        p[m] = G ** (m + 1)
    return tau, ssa, g, p


def _z(shape, device):
    return fzeros(shape, device=device)


KINDS = ["1scalar", "2stream", "nstream"]


@pytest.mark.parametrize("k1", KINDS)
@pytest.mark.parametrize("k2", KINDS)
@pytest.mark.parametrize("bybnd", [False, True])
def test_increment_with_transparent(backend, k1, k2, bybnd):
    lib, device = backend
    tau, ssa, g, p = _ref()
    d = lambda a: rc.dev(a.copy(order="F"), device)
    t1, s1, g1, p1 = d(tau), d(ssa), d(g), d(p)
    t2, s2, g2 = _z(tau.shape, device), _z(tau.shape, device), _z(tau.shape, device)
    p2 = _z(p.shape, device)
    op1 = {"1scalar": [t1], "2stream": [t1, s1, g1], "nstream": [t1, s1, p1]}[k1]
    op2 = {"1scalar": [t2], "2stream": [t2, s2, g2], "nstream": [t2, s2, p2]}[k2]
    if k1 == "1scalar" and k2 != "1scalar":
        op2 = op2[:2]  # 1scl += tau2*(1-ssa2)
    if k1 != "1scalar" and k2 == "1scalar":
        op1 = op1[:2]  # g / p unchanged
    sizes = [NCOL, NLAY, NGPT]
    if k1 == "nstream" and k2 != "1scalar":
        sizes.append(NMOM)
    if k2 == "nstream" and k1 != "1scalar":
        sizes.append(NMOM)
    if bybnd:
        lims = np.array([[1], [NGPT]], dtype=np.int32, order="F")
        name = f"rte_inc_{k1}_by_{k2}_bybnd"
        lib.call(name, *sizes, *op1, *op2, 1, rc.dev(lims, device) if device else lims)
    else:
        lib.call(f"rte_increment_{k1}_by_{k2}", *sizes, *op1, *op2)
    lib.sync()
    assert rc.allclose(rc.host(t1), tau)
    if k1 != "1scalar":
        assert rc.allclose(rc.host(s1), ssa)
    if k1 == "2stream":
        assert rc.allclose(rc.host(g1), g)
    if k1 == "nstream":
        assert rc.allclose(rc.host(p1), p)


def test_half_plus_half_is_whole(backend):
    lib, device = backend
    tau, ssa, g, p = _ref()
    d = lambda a: rc.dev(a.copy(order="F"), device)
    # 1scl
    t = d(0.5 * tau)
    t_b = d(0.5 * tau)
    lib.rte_increment_1scalar_by_1scalar(NCOL, NLAY, NGPT, t, t_b)
    lib.sync()
    assert rc.allclose(rc.host(t), tau)
    # 2str
    t, s, gg = d(0.5 * tau), d(ssa), d(g)
    t_b, s_b, g_b = d(0.5 * tau), d(ssa), d(g)
    lib.rte_increment_2stream_by_2stream(NCOL, NLAY, NGPT, t, s, gg, t_b, s_b, g_b)
    lib.sync()
    assert rc.allclose(rc.host(t), tau) and rc.allclose(rc.host(s), ssa) and rc.allclose(rc.host(gg), g)
    # nstr
    t, s, pp = d(0.5 * tau), d(ssa), d(p)
    t_b, s_b, p_b = d(0.5 * tau), d(ssa), d(p)
    lib.rte_increment_nstream_by_nstream(NCOL, NLAY, NGPT, NMOM, NMOM, t, s, pp, t_b, s_b, p_b)
    lib.sync()
    assert rc.allclose(rc.host(t), tau) and rc.allclose(rc.host(s), ssa) and rc.allclose(rc.host(pp), p)


def test_delta_scale_f0_is_identity_and_g2_formula(backend):
    lib, device = backend
    tau, ssa, g, _ = _ref()
    d = lambda a: rc.dev(a.copy(order="F"), device)
    t, s, gg = d(tau), d(ssa), d(g)
    lib.rte_delta_scale_2str_f_k(NCOL, NLAY, NGPT, t, s, gg, _z(tau.shape, device))
    lib.sync()
    assert rc.allclose(rc.host(t), tau) and rc.allclose(rc.host(s), ssa) and rc.allclose(rc.host(gg), g)
    # f = g^2 variant against the closed form (mo_optical_props_kernels.F90:89-93)
    t, s, gg = d(tau), d(ssa), d(g)
    lib.rte_delta_scale_2str_k(NCOL, NLAY, NGPT, t, s, gg)
    lib.sync()
    f = g * g
    wf = ssa * f
    assert rc.allclose(rc.host(t), (1 - wf) * tau)
    assert rc.allclose(rc.host(s), (ssa - wf) / (1 - wf))
    assert rc.allclose(rc.host(gg), (g - f) / (1 - f))


def test_extract_subsets(backend):
    lib, device = backend
    rng = np.random.default_rng(7)
    ncol, nlay, ngpt, nmom = 11, 5, 3, 2
    a = np.asfortranarray(rng.random((ncol, nlay, ngpt)))
    b = np.asfortranarray(rng.random((ncol, nlay, ngpt)))
    p = np.asfortranarray(rng.random((nmom, ncol, nlay, ngpt)))
    cs, ce = 3, 9
    n = ce - cs + 1
    out = _z((n, nlay, ngpt), device)
    lib.rte_extract_subset_dim1_3d(ncol, nlay, ngpt, rc.dev(a, device), cs, ce, out)
    lib.sync()
    assert np.array_equal(rc.host(out), a[cs - 1:ce])
    out4 = _z((nmom, n, nlay, ngpt), device)
    lib.rte_extract_subset_dim2_4d(nmom, ncol, nlay, ngpt, rc.dev(p, device), cs, ce, out4)
    lib.sync()
    assert np.array_equal(rc.host(out4), p[:, cs - 1:ce])
    lib.rte_extract_subset_absorption_tau(ncol, nlay, ngpt, rc.dev(a, device), rc.dev(b, device), cs, ce, out)
    lib.sync()
    assert rc.allclose(rc.host(out), (a * (1.0 - b))[cs - 1:ce])
