#!/usr/bin/env python
"""Per-GPU throughput of the other BASELINE.json configurations (they are parity-test cases, not bench lines;
these numbers are supplementary evidence kept under profiles/).  Synthetic inputs as in SURVEY 8(d).
  C3  RFMIP-like clear sky: 1800 distinct profiles tiled to 131,072 columns per GPU, 60 layers, top_at_1, LW+SW
  C4  all-sky, reduced 128/112 g-point k-distributions, 524,288 columns per GPU streamed in chunks of 131,072
  C5  all-sky, LW two-stream + aerosols, 131,072 columns per GPU streamed in chunks of 32,768 (g-point fluxes out)
usage: python tools/bench_configs.py [c3 c4 c5]   -> one JSON line per configuration"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rte_rrtmgp_b200 as pkg  # noqa: E402
from rte_rrtmgp_b200 import synthetic as syn  # noqa: E402
from rte_rrtmgp_b200.allsky import AllSky  # noqa: E402
from rte_rrtmgp_b200.frontend import Context  # noqa: E402


def tile(prof, n):
    reps = -(-n // next(iter(prof.values())).shape[0])
    return {k: np.asfortranarray(np.tile(v, (reps,) + (1,) * (v.ndim - 1))[:n]) for k, v in prof.items()}


def run(name, total_cols, chunk, nlay, kd_lw, kd_sw, steps=3, **kw):
    lib = pkg.lib()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx = Context(lib, "cuda:0")
    profiles = kw.pop("profiles", None)
    sky = AllSky(ctx, chunk, nlay, kd_lw, kd_sw, profiles=profiles, **kw)
    sky.step()
    ctx.config_checks(False, False)
    nchunk = total_cols // chunk
    for _ in range(2):
        sky.step()
    torch.cuda.synchronize()
    import ctypes
    lib.cdll.rrtmgpb_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        for _ in range(nchunk):  # the same resident chunk stands in for every chunk (inputs are synthetic)
            sky.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.cdll.rrtmgpb_profile_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.cdll.rrtmgpb_profile_report(buf, ctypes.c_size_t(len(buf)))
    kernels = {}
    for ln in buf.value.decode().splitlines()[:8]:
        kname, cnt, tot = ln.rsplit(" ", 2)
        kernels[kname] = round(float(tot) / steps, 2)
    f = sky.fluxes_host()
    ok = all(np.isfinite(v).all() for v in f.values())
    print(json.dumps({"config": name, "columns_per_gpu": total_cols, "chunk_columns": chunk, "nlay": nlay,
                      "ngpt_lw": kd_lw.ngpt, "ngpt_sw": kd_sw.ngpt, "ms_per_pass": ms,
                      "columns_per_s_per_gpu": total_cols / (ms * 1e-3), "finite": ok, "kernel_ms_per_pass": kernels, **{k: str(v) for k, v in kw.items()}}),
          flush=True)
    del sky
    torch.cuda.empty_cache()


def main():
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    if "c3" in which:
        prof = tile(syn.perturbed_profiles(1800, 60, seed=1234, top_at_1=True), 131072)
        run("C3 RFMIP-like clear-sky LW+SW", 131072, 131072, 60, syn.make_kdist("lw"), syn.make_kdist("sw"),
            profiles=prof, do_clouds=False)
    if "c4" in which:
        run("C4 all-sky, reduced k-distributions", 524288, 131072, 72, syn.make_kdist("lw", ngpt=128),
            syn.make_kdist("sw", ngpt=112))
    if "c5" in which:
        run("C5 all-sky, LW two-stream + aerosols", 131072, 32768, 72, syn.make_kdist("lw"), syn.make_kdist("sw"),
            do_aerosols=True, lw_2stream=True)


if __name__ == "__main__":
    main()
