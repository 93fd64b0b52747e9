#!/usr/bin/env python
"""Per-kernel times of one all-sky step (event profiler of the library), for A/B runs of build variants:
  RRTMGPB_LIB=rte_rrtmgp_b200/lib/variants/x.so python tools/kbench.py [--ncol N] [--nlay L] [--steps K] [--distinct] [--tag T]
Prints one JSON line {tag, ms_per_step, kernels: {name: ms}}.  Not a bench.py replacement (no e2e, no clocks)."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncol", type=int, default=65536)
    ap.add_argument("--nlay", type=int, default=72)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--distinct", action="store_true", help="RFMIP-like distinct columns (clear sky) instead of the replicated all-sky profile")
    ap.add_argument("--tag", default=os.environ.get("RRTMGPB_LIB", "main"))
    ap.add_argument("--express", action="store_true", help="the express path (no (ncol,nlay,ngpt) arrays)")
    ap.add_argument("--seq", action="store_true", help="the reference call sequence: the 45 extern symbols kernel by kernel")
    ap.add_argument("--rows", type=int, default=-1, help="rrtmgpb_set_gas_optics_rows_path (-1: environment)")
    ap.add_argument("--lw-only", action="store_true")
    ap.add_argument("--sw-only", action="store_true")
    a = ap.parse_args()
    import torch

    import rte_rrtmgp_b200 as pkg
    from rte_rrtmgp_b200 import synthetic as syn
    from rte_rrtmgp_b200.allsky import AllSky
    from rte_rrtmgp_b200.frontend import Context

    lib = pkg.lib()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx = Context(lib, "cuda:0")
    if a.rows >= 0:
        lib.cdll.rrtmgpb_set_gas_optics_rows_path(a.rows)
    kd_lw = None if a.sw_only else syn.make_kdist("lw")
    kd_sw = None if a.lw_only else syn.make_kdist("sw")
    prof = None
    if a.distinct:
        base = syn.perturbed_profiles(1800, a.nlay, seed=1234, top_at_1=True)
        reps = -(-a.ncol // 1800)
        import numpy as np
        prof = {k: np.asfortranarray(np.tile(v, (reps,) + (1,) * (v.ndim - 1))[:a.ncol]) for k, v in base.items()}
    free0 = torch.cuda.mem_get_info()[0]
    sky = AllSky(ctx, a.ncol, a.nlay, kd_lw, kd_sw, do_clouds=not a.distinct, profiles=prof, express=a.express)
    if a.seq:
        sky.fused = False
    sky.step()
    ctx.config_checks(False, False)
    sky.step()
    torch.cuda.synchronize()
    lib.cdll.rrtmgpb_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        sky.step()
    e1.record()
    torch.cuda.synchronize()
    lib.cdll.rrtmgpb_profile_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.cdll.rrtmgpb_profile_report(buf, ctypes.c_size_t(len(buf)))
    ks = {}
    for ln in buf.value.decode().splitlines():
        name, cnt, tot = ln.rsplit(" ", 2)
        ks[name] = round(float(tot) / a.steps, 3)
    mem = torch.cuda.max_memory_allocated() / 2**30
    used = (free0 - torch.cuda.mem_get_info()[0]) / 2**30   # includes the library's stream-ordered pool at its peak
    print(json.dumps({"tag": a.tag, "express": a.express, "peak_torch_GiB": round(mem, 2), "device_GiB": round(used, 2), "ncol": a.ncol, "nlay": a.nlay, "distinct": a.distinct,
                      "ms_per_step": round(e0.elapsed_time(e1) / a.steps, 3), "kernels": ks}), flush=True)


if __name__ == "__main__":
    main()
