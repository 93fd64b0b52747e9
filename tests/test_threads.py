"""Re-entrancy of the C-ABI (SURVEY 8b "Threading"): the reference kernels are serial and stateless, so a host may drive
disjoint column blocks from several threads at once (its OpenMP-over-blocks idiom,
examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:177-178).  Two host threads run the all-sky step on two disjoint blocks
concurrently - each on its own per-thread stream (csrc/runtime.cu) - and must reproduce, bit for bit, what the same
blocks give when run one after the other; once more with the event profiler on (its records are shared state)."""
import ctypes
import threading

import numpy as np
import pytest

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import Context


def _blocks(lib, kd_lw, kd_sw, fused):
    ctx = Context(lib, "cuda:0")
    return [AllSky(ctx, n, 60, kd_lw, kd_sw, profiles=syn.perturbed_profiles(n, 60, seed=s, top_at_1=True), fused=fused,
                   col_offset=o) for n, s, o in ((48, 1, 0), (80, 2, 48))]


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("profile", [False, True])
def test_two_host_threads_two_column_blocks(cuda_lib, fused, profile):
    kd_lw, kd_sw = syn.make_kdist("lw", ngpt=64), syn.make_kdist("sw", ngpt=56)
    serial = _blocks(cuda_lib, kd_lw, kd_sw, fused)
    for b in serial:
        b.step()
    cuda_lib.sync()
    want = [b.fluxes_host() for b in serial]

    conc = _blocks(cuda_lib, kd_lw, kd_sw, fused)
    errors = []
    streams = [None, None]

    def work(i):
        try:
            streams[i] = cuda_lib.cdll.rrtmgpb_get_stream()
            for _ in range(3):  # several steps: more chances for the two threads' launches to interleave
                conc[i].step()
            cuda_lib.sync()  # this thread's stream
        except Exception as e:  # pragma: no cover
            errors.append(e)

    cuda_lib.cdll.rrtmgpb_profile_enable(1 if profile else 0)
    try:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    finally:
        cuda_lib.cdll.rrtmgpb_profile_enable(0)
    assert not errors, errors
    if profile:
        buf = ctypes.create_string_buffer(1 << 16)
        n = cuda_lib.cdll.rrtmgpb_profile_report(buf, ctypes.c_size_t(len(buf)))
        assert n > 0 and b"sw_2stream" in buf.value
    import torch

    torch.cuda.synchronize()
    for b, w in zip(conc, want):
        got = b.fluxes_host()
        for k in w:
            assert np.array_equal(got[k], w[k]), k


@pytest.mark.gpu
def test_tables_changed_invalidates_the_gfast_copies(oracle_lib, cuda_lib):
    """A host that refills a k-distribution table IN PLACE must call rrtmgpb_tables_changed(); afterwards the fused path
    uses the new coefficients (without the call the transposed copies would be stale by design)."""
    kd_lw = syn.make_kdist("lw", ngpt=64)
    ctx = Context(cuda_lib, "cuda:0")
    sky = AllSky(ctx, 32, 60, kd_lw, None, do_clouds=False, fused=True)
    sky.step()
    tau0 = ctx.get(sky.lw.atmos.tau).copy()
    c = cuda_lib.cdll
    c.rrtmgpb_gas_optics_kmajor.restype = ctypes.c_void_p
    c.rrtmgpb_gas_optics_kmajor.argtypes = [ctypes.c_void_p]
    c.rrtmgpb_mem_to_backend.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    c.rrtmgpb_tables_changed.argtypes = [ctypes.c_void_p]
    kmajor_dev = c.rrtmgpb_gas_optics_kmajor(sky.lw.go.handle)
    doubled = np.asfortranarray(kd_lw.kmajor * 2.0)
    c.rrtmgpb_mem_to_backend(kmajor_dev, doubled.ctypes.data, doubled.nbytes)   # refill in place: same allocation
    cuda_lib.sync()
    c.rrtmgpb_tables_changed(kmajor_dev)
    sky.step()
    tau1 = ctx.get(sky.lw.atmos.tau)
    kd2 = syn.make_kdist("lw", ngpt=64)
    kd2.kmajor = np.asfortranarray(kd2.kmajor * 2.0)
    ref = AllSky(Context(oracle_lib, None), 32, 60, kd2, None, do_clouds=False)
    ref.step()
    np.testing.assert_allclose(tau1, ref.ctx.get(ref.lw.atmos.tau), rtol=1e-12)
    assert np.max(np.abs(tau1 - tau0)) > 0
    c.rrtmgpb_tables_changed(None)                   # drop everything: the next call rebuilds
    sky.step()
    np.testing.assert_array_equal(ctx.get(sky.lw.atmos.tau), tau1)
