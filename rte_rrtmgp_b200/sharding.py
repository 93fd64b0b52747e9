"""Column sharding across ranks.  No kernel has a cross-column term, so rank r of R
owns the contiguous columns [r*N/R, (r+1)*N/R) of every input and output and the hot path needs no
collective.  The only exchange is the optional gather of the broadband fluxes (ncol, nlay+1) to rank 0 -
NCCL when the tensors live on GPUs, gloo on CPU."""
import numpy as np
import torch
import torch.distributed as dist


def column_shard(ncol, rank, world):
    """[lo, hi) of rank's contiguous column range; sizes differ by at most one."""
    base, rem = divmod(ncol, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_fluxes(fluxes, ncol, rank, world, device=None):
    """Gather {name: (ncol_local, nlev) array} to rank 0 in column order; returns the dict on rank 0, None
    elsewhere.  Shards may be ragged, so every rank pads to the largest shard."""
    if world == 1:
        return fluxes
    nmax = max(column_shard(ncol, r, world)[1] - column_shard(ncol, r, world)[0] for r in range(world))
    out = {}
    for name in sorted(fluxes):
        a = np.ascontiguousarray(np.asarray(fluxes[name]))
        pad = np.zeros((nmax,) + a.shape[1:], dtype=a.dtype)
        pad[: a.shape[0]] = a
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, bufs, dst=0)
        if rank == 0:
            parts = []
            for r, b in enumerate(bufs):
                lo, hi = column_shard(ncol, r, world)
                parts.append(b[: hi - lo].cpu().numpy())
            out[name] = np.asfortranarray(np.concatenate(parts, axis=0))
    return out if rank == 0 else None


def gather_fluxes_device(tensors, rank, world, out=None):
    """Device-resident variant for equal shards: gathers each (nlev, ncol_local) torch tensor (the Fortran (ncol, nlev)
    array seen row-major) to rank 0 with one NCCL gather per array, no host staging.  `out`: optional list of per-array
    lists of receive buffers to reuse on rank 0.  Returns the receive buffers on rank 0 (column order = rank order)."""
    if world == 1:
        return [[t] for t in tensors]
    res = []
    for i, t in enumerate(tensors):
        # Fortran-ordered arrays are transposed views of contiguous storage: gather that storage
        tc = t if t.is_contiguous() else t.permute(*reversed(range(t.dim())))
        assert tc.is_contiguous()
        bufs = None
        if rank == 0:
            bufs = out[i] if out is not None else [torch.empty_like(tc) for _ in range(world)]
        dist.gather(tc, bufs, dst=0)
        res.append(bufs)
    return res if rank == 0 else None
