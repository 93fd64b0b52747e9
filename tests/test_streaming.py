"""The library's host-buffer entry (rrtmgpb_allsky_stream_host): state and fluxes in HOST memory, column chunks streamed
through the device with uploads / downloads overlapped inside the library.  Must reproduce the device-resident AllSky
driver on the same inputs - bit for bit on the plane path (same kernels on the same columns; a chunk boundary changes
nothing because columns are independent), within rounding of the broadband sums on the express path - for chunk widths
that do not divide ncol, pinned and pageable memory, distinct columns."""
import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import Context
from rte_rrtmgp_b200.streaming import HostAllSky


@pytest.mark.gpu
@pytest.mark.parametrize("express", [False, True])
@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("chunk", [64, 50, 1000])
def test_host_streaming_equals_resident_driver(oracle_lib, cuda_lib, chunk, pinned, express):
    kd_lw, kd_sw = syn.make_kdist("lw", ngpt=64), syn.make_kdist("sw", ngpt=56)
    ncol, nlay = 230, 60
    prof = syn.perturbed_profiles(ncol, nlay, seed=7, top_at_1=True)
    ref = AllSky(Context(cuda_lib, "cuda:0"), ncol, nlay, kd_lw, kd_sw, profiles=prof, fused=True)
    ref.step()
    want = ref.fluxes_host()
    h = HostAllSky(cuda_lib, ncol, nlay, kd_lw, kd_sw, chunk, profiles=prof, express=express, pinned=pinned)
    h.step()
    h.step()   # a second pass over the same buffers: the double-buffer events must leave a reusable state
    got = h.fluxes_host()
    for k in want:
        if express:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-12, atol=1e-9, err_msg=k)
        else:
            assert np.array_equal(got[k], want[k]), k
    orc = AllSky(Context(oracle_lib, None), ncol, nlay, kd_lw, kd_sw, profiles=prof)
    orc.step()
    for k, v in orc.fluxes_host().items():
        assert np.max(np.abs(got[k] - v)) <= 1.0e-5, k


@pytest.mark.gpu
def test_host_streaming_lw_only_clear_sky_and_errors(cuda_lib):
    kd_lw = syn.make_kdist("lw", ngpt=64)
    h = HostAllSky(cuda_lib, 90, 72, kd_lw, None, 32, do_clouds=False)
    h.step()
    ref = AllSky(Context(cuda_lib, "cuda:0"), 90, 72, kd_lw, None, do_clouds=False, fused=True)
    ref.step()
    for k, v in ref.fluxes_host().items():
        assert np.array_equal(h.fluxes_host()[k], v), k
    bad = HostAllSky(cuda_lib, 40, 72, kd_lw, None, 16, do_clouds=False, emis=1.5)
    with pytest.raises(RuntimeError, match="rte_lw: sfc_emis has values < 0 or > 1"):
        bad.step()


@pytest.mark.parametrize("express", [False, True])
def test_oracle_statement_of_the_driver(oracle_lib, express):
    """The CPU statement of the same driver (oracle/allsky_stream_ref.cpp) against the oracle's resident AllSky run: bit-exact."""
    kd_lw, kd_sw = syn.make_kdist("lw", ngpt=32), syn.make_kdist("sw", ngpt=28)
    ncol, nlay = 23, 30
    prof = syn.perturbed_profiles(ncol, nlay, seed=3, top_at_1=False)
    h = HostAllSky(oracle_lib, ncol, nlay, kd_lw, kd_sw, 10, profiles=prof, express=express, pinned=False, device=None)
    h.step()
    ref = AllSky(Context(oracle_lib, None), ncol, nlay, kd_lw, kd_sw, profiles=prof)
    ref.step()
    for k, v in ref.fluxes_host().items():
        assert np.array_equal(h.fluxes_host()[k], v), k
