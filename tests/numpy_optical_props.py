"""A SECOND, independent restatement of the optical-properties arithmetic (rte/kernels/mo_optical_props_kernels.F90:47-706) in
vectorised numpy, written from the Fortran: both delta scalings, the nine increments at equal spectral resolution, and the nine
by-band increments (the second operand defined per band: the Fortran repeats the same arithmetic for every g-point of the
band, :366-706).  Arrays are Fortran-shaped (ncol, nlay, ngpt) / moments first (nmom, ncol, nlay, ngpt); every expression keeps
the Fortran's association.  tests/test_oracle_crosscheck.py requires oracle/rte_optical_props_ref.c to agree BIT FOR BIT."""
import numpy as np

EPS = 3.0 * np.finfo(np.float64).tiny   # :38


def delta_scale_2str_f(tau, ssa, g, f):                      # :47-70
    wf = ssa * f
    return (1.0 - wf) * tau, (ssa - wf) / np.maximum(EPS, 1.0 - wf), (g - f) / np.maximum(EPS, 1.0 - f)


def delta_scale_2str(tau, ssa, g):                           # :75-96
    f = g * g
    wf = ssa * f
    return (1.0 - wf) * tau, (ssa - wf) / np.maximum(EPS, 1.0 - wf), (g - f) / np.maximum(EPS, 1.0 - f)


def increment(k1, k2, op1, op2):
    """op = [tau] | [tau, ssa, g] | [tau, ssa, p(nmom, ...)] at the SAME spectral resolution.  Returns the new op1 list."""
    tau1 = op1[0]
    tau2 = op2[0]
    if k1 == "1scalar":
        if k2 == "1scalar":
            return [tau1 + tau2]                                                   # :116-127
        return [tau1 + tau2 * (1.0 - op2[1])]                                      # :131-165
    ssa1 = op1[1]
    tau12 = tau1 + tau2
    if k2 == "1scalar":                                                            # :170-190, :262-280: g / p unchanged
        return [tau12, tau1 * ssa1 / np.maximum(EPS, tau12), op1[2]]
    ssa2 = op2[1]
    tauscat12 = tau1 * ssa1 + tau2 * ssa2
    den = np.maximum(EPS, tauscat12)
    if k1 == "2stream":
        second = op2[2] if k2 == "2stream" else op2[2][0]                          # :194-258: first moment of an n-stream operand
        g1 = (tau1 * ssa1 * op1[2] + tau2 * ssa2 * second) / den
        return [tau12, tauscat12 / np.maximum(EPS, tau12), g1]
    p1 = op1[2].copy()
    nmom1 = p1.shape[0]
    if k2 == "2stream":                                                            # :284-318: Henyey-Greenstein moments g, g^2, ...
        moms = [op2[2]]
        for _ in range(1, nmom1):
            moms.append(moms[-1] * op2[2])
        for m in range(nmom1):
            p1[m] = (tau1 * ssa1 * p1[m] + tau2 * ssa2 * moms[m]) / den
    else:                                                                          # :322-358: common moments only, the others untouched
        for m in range(min(nmom1, op2[2].shape[0])):
            p1[m] = (tau1 * ssa1 * p1[m] + tau2 * ssa2 * op2[2][m]) / den
    return [tau12, tauscat12 / np.maximum(EPS, tau12), p1]


def increment_bybnd(k1, k2, op1, op2, band_lims_gpt):
    """op2 per band (ncol, nlay, nbnd): expand to g-points with band_lims_gpt (2, nbnd), 1-based inclusive."""
    ngpt = op1[0].shape[-1]
    band_of = np.zeros(ngpt, dtype=int)
    for b in range(band_lims_gpt.shape[1]):
        band_of[band_lims_gpt[0, b] - 1:band_lims_gpt[1, b]] = b
    return increment(k1, k2, op1, [a[..., band_of] for a in op2])
