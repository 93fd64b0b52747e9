"""Simple-spectral-model gas-optics kernels (ssm/mo_optics_ssm_kernels.F90:29-108): oracle vs a
numpy statement of the same loops, CUDA vs oracle (1e-14 relative: nvcc may contract the products into FMAs), plus the
properties the model rests on - tau is linear in the layer masses and the broadened tau is the unbroadened one times
play/pref."""
import ctypes as C

import numpy as np
import pytest

from rte_rrtmgp_b200.frontend import Context, _addr

NCOL, NLAY, NNU = 29, 17, 41
GRAV = 9.80665


def _p(a):
    return C.c_void_p(_addr(a))


def _i(v):
    return C.byref(C.c_int(v))


def _d(v):
    return C.byref(C.c_double(v))


def _inputs(ngas, seed=2):
    rng = np.random.default_rng(seed)
    f = np.asfortranarray
    vmr = f(rng.uniform(1e-6, 2e-2, (ngas, NCOL, NLAY)))
    plev = f(np.sort(rng.uniform(50.0, 101325.0, (NCOL, NLAY + 1)), axis=1)[:, ::-1])
    play = f(0.5 * (plev[:, 1:] + plev[:, :-1]))
    mw = rng.uniform(0.016, 0.048, ngas)
    k = f(10.0 ** rng.uniform(-4, 2, (ngas, NNU)))
    return vmr, plev, play, mw, k


def _run(lib, device, ngas, pref):
    ctx = Context(lib, device)
    vmr, plev, play, mw, k = _inputs(ngas)
    d_vmr, d_plev, d_play, d_mw, d_k = (ctx.put(a) for a in (vmr, plev, play, mw, k))
    mass = ctx.zeros((ngas, NCOL, NLAY))
    tau = ctx.zeros((NCOL, NLAY, NNU))
    ctx.c.ssm_compute_layer_mass(_i(NCOL), _i(NLAY), _i(ngas), _p(d_vmr), _p(d_plev), _p(d_mw), _d(0.028964), _p(mass))
    ctx.c.ssm_compute_tau_absorption(_i(NCOL), _i(NLAY), _i(NNU), _i(ngas), _p(d_k), _p(d_play), _d(pref), _p(mass), _p(tau))
    return ctx.get(mass), ctx.get(tau)


@pytest.mark.parametrize("ngas", [1, 2, 11])
@pytest.mark.parametrize("pref", [0.0, 50000.0])
def test_oracle_is_the_reference_loop(oracle_lib, ngas, pref):
    mass, tau = _run(oracle_lib, None, ngas, pref)
    vmr, plev, play, mw, k = _inputs(ngas)
    ref_mass = vmr * (mw / 0.028964)[:, None, None] * np.abs(plev[:, 1:] - plev[:, :-1])[None] / GRAV
    np.testing.assert_array_equal(mass, ref_mass)
    s = np.zeros((NCOL, NLAY, NNU))
    for ig in range(ngas):
        s = s + mass[ig][:, :, None] * k[ig][None, None, :]
    ref_tau = s * play[:, :, None] / pref if pref > 0 else s
    np.testing.assert_array_equal(tau, ref_tau)


def test_broadening_and_linearity(oracle_lib):
    _, t0 = _run(oracle_lib, None, 2, 0.0)
    _, t1 = _run(oracle_lib, None, 2, 50000.0)
    _, _, play, _, _ = _inputs(2)
    np.testing.assert_allclose(t1, t0 * play[:, :, None] / 50000.0, rtol=1e-15)
    assert np.all(t0 > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("ngas", [1, 2, 11])
@pytest.mark.parametrize("pref", [0.0, 50000.0])
def test_cuda_vs_oracle(oracle_lib, cuda_lib, ngas, pref):
    rm, rt = _run(oracle_lib, None, ngas, pref)
    gm, gt = _run(cuda_lib, "cuda:0", ngas, pref)
    np.testing.assert_allclose(gm, rm, rtol=1e-14, atol=0)
    np.testing.assert_allclose(gt, rt, rtol=1e-14, atol=0)
