"""McICA cloud sampling (rte/extensions/mo_cloud_sampling.F90): sampled_mask_max_ran,
sampled_mask_exp_ran, draw_samples.  The oracle is checked against a line-by-line numpy transcription of the Fortran
column loop and against the properties the overlap rules imply; the CUDA kernels against the oracle, bit for bit."""
import numpy as np
import pytest

from rte_rrtmgp_b200.frontend import (Context, OpticalProps, draw_samples, sampled_mask_exp_ran, sampled_mask_max_ran)


def _ctx(kind):
    if kind == "oracle":
        import oracle

        return Context(oracle.lib(), None)
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import rte_rrtmgp_b200

    return Context(rte_rrtmgp_b200.lib(), "cuda:0")


def _inputs(ncol, nlay, ngpt, seed=3):
    rng = np.random.default_rng(seed)
    randoms = np.asfortranarray(rng.uniform(0.0, 1.0, (ngpt, nlay, ncol)))
    frac = rng.uniform(0.0, 1.0, (ncol, nlay))
    frac[rng.uniform(size=(ncol, nlay)) < 0.4] = 0.0   # clear layers inside the cloud deck
    frac[0, :] = 0.0                                   # a clear column
    frac[1, :] = 1.0                                   # an overcast column
    frac[2, : nlay // 2] = 0.0
    overlap = rng.uniform(-1.0, 1.0, (ncol, nlay - 1))
    return randoms, np.asfortranarray(frac), np.asfortranarray(overlap)


def _fortran_mask(randoms, frac, overlap=None):
    """mo_cloud_sampling.F90:160-190 / :250-290, column by column, vector over g-points."""
    ngpt, nlay, ncol = randoms.shape
    mask = np.zeros((ncol, nlay, ngpt), dtype=bool)
    for icol in range(ncol):
        layer = frac[icol, :] > 0.0
        if not layer.any():
            continue
        fst = int(np.argmax(layer))
        lst = nlay - 1 - int(np.argmax(layer[::-1]))
        local = randoms[:, fst, icol].copy()
        mask[icol, fst, :] = local > (1.0 - frac[icol, fst])
        for ilay in range(fst + 1, lst + 1):
            if layer[ilay]:
                if overlap is None:
                    if not layer[ilay - 1]:
                        local = randoms[:, ilay, icol].copy()
                elif layer[ilay - 1]:
                    rho = overlap[icol, ilay - 1]
                    local = rho * (local - 0.5) + np.sqrt(1.0 - rho * rho) * (randoms[:, ilay, icol] - 0.5) + 0.5
                else:
                    local = randoms[:, ilay, icol].copy()
                mask[icol, ilay, :] = local > (1.0 - frac[icol, ilay])
    return mask


@pytest.mark.parametrize("shape", [(9, 12, 8), (5, 7, 6), (33, 16, 13)])  # ngpt not a multiple of 4 included
def test_oracle_masks_follow_the_fortran(shape):
    ncol, nlay, ngpt = shape
    ctx = _ctx("oracle")
    randoms, frac, overlap = _inputs(ncol, nlay, ngpt)
    got = ctx.get(sampled_mask_max_ran(ctx, ctx.put(randoms), ctx.put(frac)))
    assert np.array_equal(got.astype(bool), _fortran_mask(randoms, frac))
    got = ctx.get(sampled_mask_exp_ran(ctx, ctx.put(randoms), ctx.put(frac), ctx.put(overlap)))
    assert np.array_equal(got.astype(bool), _fortran_mask(randoms, frac, overlap))


def test_overlap_properties():
    ctx = _ctx("oracle")
    ncol, nlay, ngpt = 16, 10, 64
    randoms, frac, _ = _inputs(ncol, nlay, ngpt, seed=11)
    m = ctx.get(sampled_mask_max_ran(ctx, ctx.put(randoms), ctx.put(frac))).astype(bool)
    assert not m[0].any() and m[1].all()                   # clear / overcast columns
    assert not m[frac == 0.0].any()                        # no cloud where the fraction is zero
    # maximum overlap inside a contiguous deck: the smaller fraction's cloudy g-points are a subset of the larger's
    for icol in range(ncol):
        for l in range(1, nlay):
            if frac[icol, l] > 0 and frac[icol, l - 1] > 0:
                lo, hi = (l, l - 1) if frac[icol, l] <= frac[icol, l - 1] else (l - 1, l)
                assert not (m[icol, lo] & ~m[icol, hi]).any()
    # rho = 1 in exponential-random overlap reproduces maximum-random; rho = 0 uses fresh deviates in every layer
    ones = np.ones((ncol, nlay - 1), order="F")
    e1 = ctx.get(sampled_mask_exp_ran(ctx, ctx.put(randoms), ctx.put(frac), ctx.put(ones))).astype(bool)
    assert np.array_equal(e1, m)
    e0 = ctx.get(sampled_mask_exp_ran(ctx, ctx.put(randoms), ctx.put(frac), ctx.put(0.0 * ones))).astype(bool)
    want = (np.transpose(randoms, (2, 1, 0)) > (1.0 - frac)[:, :, None]) & (frac > 0)[:, :, None]
    assert np.array_equal(e0, want)


def test_error_strings():
    ctx = _ctx("oracle")
    randoms, frac, overlap = _inputs(6, 5, 4)
    with pytest.raises(RuntimeError, match="cloud_frac\\(ncol,nlay\\) are inconsistent"):
        sampled_mask_max_ran(ctx, ctx.put(randoms), ctx.put(frac[:, :4]))
    with pytest.raises(RuntimeError, match="overlap_param\\(ncol,nlay-1\\) are inconsistent"):
        sampled_mask_exp_ran(ctx, ctx.put(randoms), ctx.put(frac), ctx.put(frac))
    bad = frac.copy(); bad[3, 2] = 1.5
    with pytest.raises(RuntimeError, match="cloud fraction values out of range"):
        sampled_mask_max_ran(ctx, ctx.put(randoms), ctx.put(bad))
    bad = overlap.copy(); bad[1, 1] = -1.5
    with pytest.raises(RuntimeError, match="overlap_param values out of range"):
        sampled_mask_exp_ran(ctx, ctx.put(randoms), ctx.put(frac), ctx.put(bad))


def _clouds(ctx, kind, ncol, nlay, lims, by_band, seed=5):
    rng = np.random.default_rng(seed)
    nb = lims.shape[1]
    blims = np.stack([np.arange(1, nb + 1), np.arange(1, nb + 1)]).astype(np.int32) if by_band else lims
    op = OpticalProps(ctx, kind, ncol, nlay, blims)
    if by_band:
        op.tau = ctx.put(rng.uniform(0.1, 5.0, (ncol, nlay, nb)))
        if kind == "2str":
            op.ssa = ctx.put(rng.uniform(0.1, 1.0, (ncol, nlay, nb)))
            op.g = ctx.put(rng.uniform(0.1, 0.9, (ncol, nlay, nb)))
    return op


@pytest.mark.parametrize("kind", ["1scl", "2str"])
def test_draw_samples_oracle(kind):
    ctx = _ctx("oracle")
    ncol, nlay = 7, 6
    lims = np.array([[1, 4, 6], [3, 5, 9]], dtype=np.int32)  # ragged bands
    ngpt = 9
    randoms, frac, _ = _inputs(ncol, nlay, ngpt)
    mask = sampled_mask_max_ran(ctx, ctx.put(randoms), ctx.put(frac))
    clouds = _clouds(ctx, kind, ncol, nlay, lims, True)
    sampled = _clouds(ctx, kind, ncol, nlay, lims, False)
    draw_samples(ctx, mask, clouds, sampled)
    m = ctx.get(mask).astype(bool)
    band = np.repeat(np.arange(3), [3, 2, 4])
    for name in (("tau",) if kind == "1scl" else ("tau", "ssa", "g")):
        want = np.where(m, ctx.get(getattr(clouds, name))[:, :, band], 0.0)
        assert np.array_equal(ctx.get(getattr(sampled, name)), want)
    with pytest.raises(RuntimeError, match="need to be the same variable type"):
        draw_samples(ctx, mask, clouds, _clouds(ctx, "2str" if kind == "1scl" else "1scl", ncol, nlay, lims, False))
    with pytest.raises(RuntimeError, match="different ncol, nlay and/or ngpt"):
        draw_samples(ctx, mask[:, :, :8], clouds, sampled)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(130, 72, 32), (37, 11, 13)])
def test_cuda_matches_oracle_bit_for_bit(shape):
    ncol, nlay, ngpt = shape
    randoms, frac, overlap = _inputs(ncol, nlay, ngpt, seed=21)
    lims = np.array([[1, ngpt // 2 + 1], [ngpt // 2, ngpt]], dtype=np.int32)
    out = {}
    for kind in ("oracle", "cuda"):
        ctx = _ctx(kind)
        m1 = sampled_mask_max_ran(ctx, ctx.put(randoms), ctx.put(frac))
        m2 = sampled_mask_exp_ran(ctx, ctx.put(randoms), ctx.put(frac), ctx.put(overlap))
        clouds = _clouds(ctx, "2str", ncol, nlay, lims, True)
        sampled = _clouds(ctx, "2str", ncol, nlay, lims, False)
        draw_samples(ctx, m2, clouds, sampled)
        out[kind] = [ctx.get(m1).astype(bool), ctx.get(m2).astype(bool), ctx.get(sampled.tau), ctx.get(sampled.ssa),
                     ctx.get(sampled.g)]
    for a, b in zip(out["cuda"], out["oracle"]):
        assert np.array_equal(a, b)
