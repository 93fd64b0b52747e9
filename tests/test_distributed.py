"""Multi-GPU path: columns are independent, so ranks own contiguous column shards
and there is NO collective on the data path; the only collective is the optional gather of broadband
fluxes to rank 0.  Covered here with world_size-2 gloo on CPU (oracle backend): sharded results gathered
to rank 0 must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rte_rrtmgp_b200 import synthetic as syn
from rte_rrtmgp_b200.allsky import AllSky
from rte_rrtmgp_b200.frontend import Context
from rte_rrtmgp_b200.sharding import column_shard, gather_fluxes, gather_fluxes_device

NCOL, NLAY = 22, 24


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle

    kd_lw, kd_sw = syn.make_kdist("lw", gpt_per_band=2), syn.make_kdist("sw", gpt_per_band=2)
    lo, hi = column_shard(NCOL, rank, world)
    prof = syn.perturbed_profiles(NCOL, NLAY, seed=7, top_at_1=False)
    mine = {k: np.asfortranarray(v[lo:hi]) for k, v in prof.items()}
    sky = AllSky(Context(oracle.lib(), None), hi - lo, NLAY, kd_lw, kd_sw, profiles=mine, col_offset=lo)
    sky.step()
    local = sky.fluxes_host()
    gathered = gather_fluxes(local, NCOL, rank, world)
    # the device-resident variant bench.py uses for N > 1 (here on CPU tensors over gloo): Fortran-ordered arrays
    # as transposed views of contiguous storage, equal shards
    names = sorted(local)
    tens = [torch.from_numpy(np.ascontiguousarray(local[k].T)).T for k in names]
    bufs = gather_fluxes_device(tens, rank, world)
    if rank == 0:
        for k, per_rank in zip(names, bufs):
            glued = np.concatenate([b.numpy().T for b in per_rank], axis=0)
            assert np.array_equal(glued, gathered[k]), k
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_column_sharding_matches_single_process():
    import oracle

    kd_lw, kd_sw = syn.make_kdist("lw", gpt_per_band=2), syn.make_kdist("sw", gpt_per_band=2)
    prof = syn.perturbed_profiles(NCOL, NLAY, seed=7, top_at_1=False)
    ref = AllSky(Context(oracle.lib(), None), NCOL, NLAY, kd_lw, kd_sw, profiles=prof)
    ref.step()
    want = ref.fluxes_host()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)


def test_column_shard_covers_range_exactly():
    for n in (1, 7, 22, 65536):
        for w in (1, 2, 3, 8):
            spans = [column_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
