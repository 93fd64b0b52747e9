#!/usr/bin/env python
"""bench.py - columns/sec of the all-sky LW+SW hot path (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5]     product (CUDA) arm
  python bench.py --impl reference [--steps K] [--warmup W]                      reference arm: the reference's CPU
                                                                                 kernels (C restatement, all host threads)

Headline workload (config.workload): BASELINE.json configs[1] = "c2" - one analytic RCEMIP-like profile replicated to
65,536 columns x 72 layers per GPU, LW 256 + SW 224 g-points, clouds in 2/3 of the columns: the loop body of the
reference's own benchmark driver (examples/all-sky/rrtmgp_allsky.F90:332-409).  A "step" = one LW iteration + one SW
iteration over all columns of the rank.  Columns are independent, so ranks own disjoint column shards and there is NO
collective on the data path ("scaling": "weak").

  value   device-timed, inputs resident in HBM, the API-visible (ncol,nlay,ngpt) arrays materialised (plane path)
  e2e     the same step through the library's HOST-buffer entry (rrtmgpb_allsky_stream_host): state and fluxes in pinned
          host memory, column chunks, uploads / downloads overlapped with compute inside the library
  extras  value_express (no (ncol,nlay,ngpt) arrays), value_reference_call_sequence (the 45 extern symbols, kernel by
          kernel), value_with_flux_gather_to_rank0 (N > 1), other_configs: c3 / c4 / c5 of BASELINE.json per GPU
          (c4 streamed from host memory through the same host-buffer entry)

--config c3|c4|c5 makes that configuration the timed headline instead (parity-test shapes; supplementary).
One JSON line is printed by rank 0 (contract: task statement / DESIGN.md section 6).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCOL_PER_GPU = 65536
NLAY = 72
METRIC = "columns/sec (LW+SW all-sky)"
WORKLOADS = {
    "c2": "all-sky LW(256 gpt)+SW(224 gpt), 1 RCEMIP-like profile replicated, 72 layers, clouds in 2/3 of columns",
    "c3": "RFMIP-like clear-sky LW(256)+SW(224), 1800 distinct profiles tiled to 131,072 columns per GPU, 60 layers",
    "c4": "all-sky LW(128 gpt)+SW(112 gpt), reduced k-distributions, 524,288 columns per GPU streamed from host memory in chunks, 72 layers",
    "c5": "all-sky LW two-stream (256 gpt, g-point fluxes summed) + SW(224), clouds + aerosols, 131,072 columns per GPU in chunks of 32,768, 72 layers",
}
# roofline columns/s per B200 from the unfused-ABI algorithmic bytes at the measured HBM bandwidth
ROOFLINE_COLS = {"c2": 1.62e6, "c3": 2.95e6, "c4": 2.71e6, "c5": 1.61e6}


def algorithmic_bytes(ncol, nlay, ngpt_lw, ngpt_sw, nbnd_lw, nbnd_sw, nflav_lw, nflav_sw, ngas):
    """SURVEY.md section 8(d) per-kernel algorithmic bytes (fp64), evaluated for the actual table sizes.
    Returns {kernel name: bytes per step} for the kernels the event profiler names."""
    w, i4, b1 = 8, 4, 1
    N, L, S = ncol, nlay, ngas + 1
    out = {}

    def interp(F):
        return (2 + S) * N * L * w + N * L * (2 * i4 + b1 + F * (2 * i4 + 14 * w))

    def planes(G):
        return N * L * G * w

    for tag, G, B, F in (("lw", ngpt_lw, nbnd_lw, nflav_lw), ("sw", ngpt_sw, nbnd_sw, nflav_sw)):
        P = planes(G)
        out[f"interpolation[{tag}]"] = interp(F)
        out[f"tau_absorption[{tag}]"] = interp(F) - (2 + S) * N * L * w + (2 + S) * N * L * w + P
        out[f"cld_from_table[{tag}]"] = 2 * 3 * N * L * B * w
    P = planes(ngpt_lw)
    out["planck_source"] = N * L * (8 * nflav_lw * w + 2 * nflav_lw * i4 + 2 * i4 + b1 + 2 * w) + P * (2 + 1.0 / L) + 2 * N * ngpt_lw * w
    out["lw_noscat_kernel"] = P * (3 + 1.0 / L) + 4 * N * ngpt_lw * w + 2 * N * (L + 1) * w
    out["lw_noscat_reg_kernel"] = out["lw_noscat_kernel"]
    out["rte_inc_1scalar_by_1scalar_bybnd"] = 2 * P
    # fused gas optics: charged only what the API makes externally visible
    small_in = N * L * (4 + ngas) * w  # play, plev, tlay, vmr
    out["gas_tau_fused[lw]"] = small_in + N * L * nbnd_lw * w + P
    out["planck_fused"] = small_in + P * (2 + 1.0 / L) + 2 * N * ngpt_lw * w
    out["gas_tau_fused[sw]"] = small_in + 3 * N * L * nbnd_sw * w + 3 * planes(ngpt_sw)
    P = planes(ngpt_sw)
    out["tau_rayleigh"] = N * L * (4 * nflav_sw * w + 2 * nflav_sw * i4 + i4 + b1 + (1 + S) * w) + P
    out["rrtmgpb_combine_abs_and_rayleigh"] = 5 * P
    out["sw_2stream_kernel"] = 3 * P + N * L * w + 4 * N * ngpt_sw * w + 3 * N * (L + 1) * w
    out["sw_2stream_reg_kernel"] = out["sw_2stream_kernel"]
    out["rte_inc_2stream_by_2stream_bybnd"] = 6 * P
    out["rte_delta_scale_2str_k"] = 6 * N * L * nbnd_sw * w
    return out


def clock_sampler(stop, samples):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    try:
        while not stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            samples.append(line.strip())
    finally:
        p.terminate()


def summarize_clocks(samples, device_index):
    sm, mx, reasons = [], [], set()
    for s in samples:
        f = [x.strip() for x in s.split(",")]
        if len(f) < 9 or f[0] != str(device_index):
            continue
        try:
            sm.append(float(f[1])); mx.append(float(f[2]))
        except ValueError:
            continue
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's default kernels) on the host cores
# ----------------------------------------------------------------------------------------------------
def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_rate(ncol_per_block, nblocks_per_thread, timed, warmup, fast=True, threads=None, lw_only_clear=False, nlay=NLAY,
             distinct=False):
    """columns/s of the CPU oracle: independent column blocks per thread - the reference's own parallelisation idiom
    (examples/rfmip-clear-sky/rrtmgp_rfmip_lw.F90:177-178,247).  Returns (median, best) columns/s over `timed`
    iterations after `warmup`, plus threads and columns per iteration."""
    import oracle
    from rte_rrtmgp_b200 import synthetic as syn
    from rte_rrtmgp_b200.allsky import AllSky
    from rte_rrtmgp_b200.frontend import Context

    threads = threads or (os.cpu_count() or 1)
    lib = oracle.lib(fast=fast)
    ctx = Context(lib, None)
    kd_lw = syn.make_kdist("lw")
    kd_sw = None if lw_only_clear else syn.make_kdist("sw")
    blocks = []
    for i in range(threads):
        prof = syn.perturbed_profiles(ncol_per_block, nlay, seed=1234 + i, top_at_1=True) if distinct else None
        blocks.append(AllSky(ctx, ncol_per_block, nlay, kd_lw, kd_sw, col_offset=i * ncol_per_block, profiles=prof,
                             do_clouds=not lw_only_clear))
    ctx.config_checks(False, False)  # rrtmgp_allsky.F90:334

    def work(b):
        for _ in range(nblocks_per_thread):
            b.step()

    def run_once():
        ts = [threading.Thread(target=work, args=(b,)) for b in blocks]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        run_once()
    times = [run_once() for _ in range(timed)]
    cols = threads * nblocks_per_thread * ncol_per_block
    return cols / float(np.median(times)), cols / min(times), threads, cols


def cpu_baseline_block():
    """single-thread and all-thread figures, parity build and speed build, C1 in full
    (1800 x 60 x 256 LW clear sky) and a slice of C2; >= 3 warm-up + 5 timed iterations, median and best."""
    ncpu = os.cpu_count() or 1
    blk = 64
    per_thread = max(1, 4096 // (blk * ncpu))
    med, best, th, cols = cpu_rate(blk, per_thread, 5, 3, fast=True)
    out = {"value": med, "unit": "columns/s", "cores": th, "kind": "port", "cpu_model": _cpu_model(),
           "sample": f"{cols} columns per iteration ({th} threads x {per_thread} blocks x {blk} columns) of the headline workload; "
                     "3 warm-up + 5 timed iterations, median (best in value_best); oracle -O3 -march=native build of the C "
                     "restatement of the reference's default kernels (no Fortran compiler in the image)",
           "value_best": best}
    m1, b1, _, c1 = cpu_rate(blk, 2, 5, 3, fast=True, threads=1)
    out["single_thread"] = {"value": m1, "value_best": b1, "columns_per_iteration": c1}
    mp, bp, _, cp = cpu_rate(blk, per_thread, 3, 1, fast=False)
    out["parity_build_all_threads"] = {"value": mp, "value_best": bp, "columns_per_iteration": cp,
                                       "flags": "-O2 -ffp-contract=off (the oracle the parity tests use)"}
    # C1: RFMIP clear-sky LW, 1800 columns x 60 layers x 256 g-points in full, blocks over threads
    nb = -(-1800 // ncpu)
    mc, bc, _, cc = cpu_rate(nb, 1, 5, 3, fast=True, lw_only_clear=True, nlay=60, distinct=True)
    m1c, b1c, _, c1c = cpu_rate(1800, 1, 3, 1, fast=True, threads=1, lw_only_clear=True, nlay=60, distinct=True)
    out["c1_rfmip_clear_sky_lw"] = {"all_threads": {"value": mc, "value_best": bc, "columns_per_iteration": cc, "threads": ncpu},
                                    "single_thread": {"value": m1c, "value_best": b1c, "columns_per_iteration": c1c},
                                    "shape": "1800 distinct columns x 60 layers x 256 g-points, LW clear sky"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ncb, nbt = 64, 2
    med, best, threads, cols = cpu_rate(ncb, nbt, max(args.steps, 1), max(args.warmup, 1))
    rate = med
    sample = (f"{cols} columns per step ({threads} threads x {nbt} blocks x {ncb} columns) of the same workload; "
              "median over the timed steps; oracle -O3 -march=native build")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "columns/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cols / rate * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["c2"], "ncol_per_step": cols, "nlay": NLAY},
        "cpu_baseline": {"value": rate, "unit": "columns/s", "cores": threads, "kind": "port", "sample": sample,
                         "value_best": best, "cpu_model": _cpu_model()},
        "e2e": {"value": rate, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import rte_rrtmgp_b200 as pkg

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product arm)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.lib = pkg.lib()
        self.lib.set_device(self.local)
        self.lib.set_stream(torch.cuda.current_stream().cuda_stream)
        self.device = f"cuda:{self.local}"

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """CUDA events around `steps` calls of fn on torch's current stream (= the library's stream), barriers on both
        sides, max over ranks; ms per step."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps

    def timed_host(self, fn, steps):
        """Wall clock around host-synchronous calls (the host-buffer entry returns when the fluxes are in host memory),
        device-synchronised and barriered on both sides, max over ranks; ms per step."""
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        self.barrier()
        return self.max_over_ranks(ms) / steps

    def profile(self, fn, steps):
        lib = self.lib
        self.torch.cuda.synchronize()
        lib.cdll.rrtmgpb_profile_enable(1)
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        lib.cdll.rrtmgpb_profile_enable(0)
        buf = ctypes.create_string_buffer(1 << 16)
        lib.cdll.rrtmgpb_profile_report(buf, ctypes.c_size_t(len(buf)))
        prof = []
        for ln in buf.value.decode().splitlines():
            name, cnt, tot = ln.rsplit(" ", 2)
            prof.append((name, int(cnt) / steps, float(tot) / steps))
        return prof

    def free(self):
        import gc

        gc.collect()
        self.torch.cuda.empty_cache()


def _tile(prof, n):
    reps = -(-n // next(iter(prof.values())).shape[0])
    return {k: np.asfortranarray(np.tile(v, (reps,) + (1,) * (v.ndim - 1))[:n]) for k, v in prof.items()}


def make_config(b, name, ncol=None):
    """-> (step function, columns per step, description dict, cleanup objects) for one BASELINE.json configuration."""
    from rte_rrtmgp_b200 import synthetic as syn
    from rte_rrtmgp_b200.allsky import AllSky
    from rte_rrtmgp_b200.frontend import Context
    from rte_rrtmgp_b200.streaming import HostAllSky

    ctx = Context(b.lib, b.device)
    if name == "c3":
        n = ncol or 131072
        prof = _tile(syn.perturbed_profiles(1800, 60, seed=1234, top_at_1=True), n)
        sky = AllSky(ctx, n, 60, syn.make_kdist("lw"), syn.make_kdist("sw"), profiles=prof, do_clouds=False, col_offset=b.rank * n)

        # (unrelated neighbouring columns: from the second step on the library picks the tau kernels' lanes-along-g-points
        #  mapping by itself - rrtmgpb_set_gas_optics_rows_path, include/rrtmgp_b200_ext.h; same results either way)
        return sky.step, n, {"ncol_per_gpu": n, "nlay": 60, "resident": True, "gas_optics_rows_path": "automatic"}, sky
    if name == "c4":
        n, chunk = ncol or 524288, 37888   # 8 solver waves per chunk
        h = HostAllSky(b.lib, n, NLAY, syn.make_kdist("lw", ngpt=128), syn.make_kdist("sw", ngpt=112), chunk, device=b.device)
        return h.step, n, {"ncol_per_gpu": n, "nlay": NLAY, "chunk_columns": chunk, "resident": False,
                           "h2d_bytes_per_step": h.h2d_bytes, "d2h_bytes_per_step": h.d2h_bytes,
                           "note": "state and fluxes in pinned host memory; every chunk's inputs cross PCIe (rrtmgpb_allsky_stream_host)"}, h
    if name == "c5":
        n, chunk = ncol or 131072, 32768
        sky = AllSky(ctx, chunk, NLAY, syn.make_kdist("lw"), syn.make_kdist("sw"), do_aerosols=True, lw_2stream=True,
                     col_offset=b.rank * n)
        nchunk = n // chunk

        def step():
            for _ in range(nchunk):   # device-resident chunk (synthetic replicated profile: every chunk is the same columns)
                sky.step()
        return step, n, {"ncol_per_gpu": n, "nlay": NLAY, "chunk_columns": chunk, "resident": True,
                         "note": "LW two-stream returns g-point fluxes (2 x 9.8 GB per chunk), summed with rte_sum_broadband"}, sky
    raise ValueError(name)


def run_product(args):
    b = Bench(args)
    torch, lib, world, rank = b.torch, b.lib, b.world, b.rank
    from rte_rrtmgp_b200 import synthetic as syn
    from rte_rrtmgp_b200.allsky import AllSky
    from rte_rrtmgp_b200.frontend import Context
    from rte_rrtmgp_b200.streaming import HostAllSky

    ctx = Context(lib, b.device)
    steps, warmup = args.steps, args.warmup
    stop, samples = threading.Event(), []
    sampler = threading.Thread(target=clock_sampler, args=(stop, samples), daemon=True)
    extras, kernels, roofline, step_roof = {}, [], None, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"

    if args.config == "c2":
        ncol = args.ncol
        kd_lw, kd_sw = syn.make_kdist("lw"), syn.make_kdist("sw")
        free0 = torch.cuda.mem_get_info()[0]
        sky = AllSky(ctx, ncol, NLAY, kd_lw, kd_sw, col_offset=rank * ncol)
        # first step with the frontend's checks on (as the reference driver does), then off: rrtmgp_allsky.F90:334
        sky.step()
        ctx.config_checks(False, False)
        for _ in range(max(warmup - 1, 0)):
            sky.step()
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        plane_gib = round((free0 - torch.cuda.mem_get_info()[0]) / 2**30, 2)   # device memory of the resident plane path
        lib.launch_count(reset=True)
        ms_per_step = b.timed(sky.step, steps)        # headline: event profiler OFF
        launches = lib.launch_count()
        value = world * ncol / (ms_per_step * 1e-3)
        workload, cfg = WORKLOADS["c2"], {"ncol_per_gpu": ncol, "nlay": NLAY, "ngpt_lw": kd_lw.ngpt, "ngpt_sw": kd_sw.ngpt,
                                          "device_GiB": plane_gib}

        # ---- per-kernel shares (separate profiled pass) and the roofline of the dominant kernel
        prof = b.profile(sky.step, max(steps // 2, 2))
        alg = algorithmic_bytes(ncol, NLAY, kd_lw.ngpt, kd_sw.ngpt, kd_lw.nbnd, kd_sw.nbnd, kd_lw.nflav, kd_sw.nflav, kd_lw.ngas)
        total_kernel_ms = sum(p[2] for p in prof) or 1.0
        ncu = {}
        for fn in ("r2_ncu_summary.json", "r1_ncu_summary.json"):
            try:
                ncu = json.load(open(os.path.join(ROOT, "profiles", fn)))
                ncu_src = fn
                break
            except Exception:
                continue
        for name, cnt, ms in prof:
            bytes_step = alg.get(name) or ((alg[f"{name}[lw]"] + alg[f"{name}[sw]"]) if f"{name}[lw]" in alg else None)
            ent = {"kernel": name, "launches_per_step": cnt, "ms_per_step": ms, "share": ms / total_kernel_ms}
            if bytes_step:
                ent["algorithmic_bytes_per_step"] = bytes_step
                ent["achieved_gbs"] = bytes_step / (ms * 1e-3) / 1e9
                ent["frac"] = ent["achieved_gbs"] / peak
            rec = ncu.get(name)
            if rec:
                ent["traffic"] = rec["dram_bytes_per_column"] * ncol * cnt / max(rec.get("launches", 1), 1)
                ent["fp64_pipe_active_pct_ncu"] = rec.get("fp64_pipe_active_pct")
            kernels.append(ent)
        if kernels:
            top = kernels[0]
            roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top.get("achieved_gbs"), "peak": peak,
                        "unit": "GB/s", "frac": top.get("frac"), "traffic": top.get("traffic"), "peak_source": peak_src,
                        "share_of_step": top["share"], "ms_per_launch": top["ms_per_step"] / max(top["launches_per_step"], 1),
                        "launches_per_step": top["launches_per_step"],
                        "algorithmic_bytes_per_launch": (top.get("algorithmic_bytes_per_step") or 0) / max(top["launches_per_step"], 1),
                        # the solvers are fp64-issue bound, not HBM bound: pipe utilisation from ncu
                        "fp64_pipe_active_pct_ncu": top.get("fp64_pipe_active_pct_ncu"),
                        "traffic_source": f"profiles/{ncu_src} (ncu --set full, dram__bytes_read+write, scaled to this launch)" if ncu else None}
        step_bytes = sum(k.get("algorithmic_bytes_per_step") or 0 for k in kernels)
        step_roof = {"algorithmic_bytes_per_step": step_bytes, "frac_of_hbm_peak": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                     "bytes_per_column": step_bytes / ncol, "frac_of_unfused_abi_roofline": value / world / ROOFLINE_COLS["c2"]}

        # ---- N > 1: the optional epilogue of SURVEY 8(e) - broadband fluxes gathered to rank 0 over NCCL after every step
        if world > 1:
            from rte_rrtmgp_b200.sharding import gather_fluxes_device
            flux_t = [sky.lw.flux_up, sky.lw.flux_dn, sky.sw.flux_up, sky.sw.flux_dn, sky.sw.flux_dir]
            recv = ([[torch.empty(tuple(reversed(t.shape)), dtype=t.dtype, device=t.device) for _ in range(world)] for t in flux_t]
                    if rank == 0 else None)

            def step_gather():
                sky.step()
                gather_fluxes_device(flux_t, rank, world, recv)
            step_gather()
            extras["value_with_flux_gather_to_rank0"] = world * ncol / (b.timed(step_gather, steps) * 1e-3)

        # ---- the same step driven as the reference's call sequence, kernel by kernel through the 45 extern-ABI symbols
        if not args.no_seq:
            sky.fused = False
            sky.step()
            extras["value_reference_call_sequence"] = world * ncol / (b.timed(sky.step, max(steps // 2, 1)) * 1e-3)
            sky.fused = True
            sky.step()  # leave the headline path's results in the flux arrays for the parity spot check below
        flux_gpu = sky.fluxes_host() if rank == 0 else None
        del sky
        b.free()

        # ---- express path: no (ncol,nlay,ngpt) arrays
        if not args.no_extras:
            free_before = torch.cuda.mem_get_info()[0]
            xs = AllSky(ctx, ncol, NLAY, kd_lw, kd_sw, col_offset=rank * ncol, express=True)
            xs.step(); xs.step()
            ms_x = b.timed(xs.step, steps)
            extras["value_express"] = world * ncol / (ms_x * 1e-3)
            # device memory this leg added (the library's pool keeps its peak): state, by-band cloud arrays, fluxes, scratch
            extras["express"] = {"ms_per_step": ms_x, "device_GiB_added": round((free_before - torch.cuda.mem_get_info()[0]) / 2**30, 2),
                                 "device_GiB_plane_path": plane_gib,
                                 "note": "rrtmgpb_rte_lw_express / _sw_express: column chunks x band groups, planes only in a reused scratch"}
            if rank == 0 and flux_gpu is not None:
                fx = xs.fluxes_host()
                extras["express"]["max_abs_flux_diff_vs_plane_path_Wm2"] = max(float(np.max(np.abs(fx[k] - flux_gpu[k]))) for k in fx)
            del xs
            b.free()

        # ---- end to end through the library's HOST-buffer entry: state and fluxes in pinned host memory
        e2e = None
        if not args.no_e2e:
            h = HostAllSky(lib, ncol, NLAY, kd_lw, kd_sw, args.e2e_chunk, device=b.device)
            h.step(); h.step()
            ms_e2e = b.timed(h.step, steps)   # events on the library's compute stream; the call returns when the fluxes are on the host
            e2e = {"value": world * ncol / (ms_e2e * 1e-3), "unit": "columns/s", "h2d_bytes_per_step": h.h2d_bytes,
                   "d2h_bytes_per_step": h.d2h_bytes, "ms_per_step": ms_e2e,
                   "chunk_columns": args.e2e_chunk or "library default (4 solver waves, small first chunk, 2 compute streams)",
                   "api": "rrtmgpb_allsky_stream_host (include/rrtmgp_b200_frontend.h): host buffers in, host fluxes out, copies inside the timed region"}
            if rank == 0 and flux_gpu is not None:
                fh = h.fluxes_host()
                e2e["max_abs_flux_diff_vs_resident_Wm2"] = max(float(np.max(np.abs(fh[k] - flux_gpu[k]))) for k in fh)
            del h
            b.free()

        # ---- a stock host (gfortran-style: HOST arrays into the 45 symbols, every call staged through PCIe), small slice;
        # in a child process: a failure on this correctness path must not take the bench line with it
        if not args.no_extras and world == 1 and rank == 0:
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_pointer_bench.py")], capture_output=True,
                                   text=True, timeout=300)
                hp = json.loads(r.stdout.strip().splitlines()[-1])
                extras["value_host_pointer_call_sequence"] = hp["value"]
                extras["host_pointer_note"] = hp["note"]
            except Exception as e:  # pragma: no cover
                extras["value_host_pointer_call_sequence"] = None
                extras["host_pointer_note"] = f"failed: {str(e)[:160]}"

        # ---- the single-precision build (RTE_ENABLE_SP): its two solvers at the headline shape, in a child process
        if not args.no_extras and world == 1 and rank == 0:
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sp_solver_bench.py"), str(ncol)], capture_output=True,
                                   text=True, timeout=300)
                extras["single_precision_solvers"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as e:  # pragma: no cover
                extras["single_precision_solvers"] = {"error": str(e)[:160]}

        # ---- the other BASELINE.json configurations, per GPU, few steps (supplementary; full runs: --config c3|c4|c5)
        if not args.no_extras:
            other = {}
            for name in ("c3", "c4", "c5"):
                try:
                    fn, n, desc, keep = make_config(b, name)
                    fn()
                    ctx.config_checks(False, False)
                    fn()
                    ms = b.timed(fn, 2)
                    v = world * n / (ms * 1e-3)
                    other[name] = {"value": v, "unit": "columns/s", "ms_per_step": ms, "workload": WORKLOADS[name],
                                   "frac_of_unfused_abi_roofline": v / world / ROOFLINE_COLS[name], **desc}
                    del fn, keep
                except Exception as e:  # pragma: no cover
                    other[name] = {"error": str(e)[:200]}
                b.free()
            extras["other_configs"] = other
    else:
        # ---- a non-headline configuration as the timed workload
        fn, ncol, desc, keep = make_config(b, args.config, None if args.ncol == NCOL_PER_GPU else args.ncol)
        fn()
        ctx.config_checks(False, False)
        for _ in range(max(warmup - 1, 0)):
            fn()
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        lib.launch_count(reset=True)
        ms_per_step = b.timed(fn, steps)
        launches = lib.launch_count()
        value = world * ncol / (ms_per_step * 1e-3)
        workload, cfg = WORKLOADS[args.config], desc
        prof = b.profile(fn, 1)
        tot = sum(p[2] for p in prof) or 1.0
        kernels = [{"kernel": n_, "launches_per_step": c_, "ms_per_step": m_, "share": m_ / tot} for n_, c_, m_ in prof]
        step_roof = {"frac_of_unfused_abi_roofline": value / world / ROOFLINE_COLS[args.config]}
        e2e = ({"value": value, "unit": "columns/s", "h2d_bytes_per_step": desc["h2d_bytes_per_step"],
                "d2h_bytes_per_step": desc["d2h_bytes_per_step"], "note": "this configuration IS the host-buffer path"}
               if args.config == "c4" else None)
        flux_gpu, kd_lw, kd_sw = None, None, None

    stop.set()
    if rank == 0:
        if sampler.is_alive():
            sampler.join(timeout=2)
        cpu_base, parity = None, None
        if world == 1 and not args.no_cpu and args.config == "c2":
            cpu_base = cpu_baseline_block()
            import oracle
            chk = AllSky(Context(oracle.lib(), None), 48, NLAY, kd_lw, kd_sw)
            chk.step()
            fc = chk.fluxes_host()
            parity = max(float(np.max(np.abs(flux_gpu[k][:48] - fc[k]))) for k in fc)
        line = {
            "metric": METRIC, "value": value, "unit": "columns/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, **cfg, "sharding": f"columns x{world}, no data-path collective",
                       "l2_policy": "inputs larger than L2 (each (col,lay,gpt) plane is GBs)"},
            "e2e": e2e,
            "gpu_launches": launches,
            **{k: extras[k] for k in ("value_with_flux_gather_to_rank0", "value_express", "value_reference_call_sequence",
                                      "value_host_pointer_call_sequence") if k in extras},
            "roofline": roofline,
            "step_roofline": step_roof,
            "clocks": summarize_clocks(samples, b.local),
            "cpu_baseline": cpu_base,
            "max_abs_flux_err_vs_oracle_Wm2": parity,
            **{k: v for k, v in extras.items() if k in ("express", "other_configs", "host_pointer_note", "single_precision_solvers")},
            "kernels": kernels[:14],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        b.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"], help="BASELINE.json configuration timed as the headline (default c2)")
    ap.add_argument("--ncol", type=int, default=NCOL_PER_GPU, help="columns per GPU (default: the BASELINE config)")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="nominal column chunk of the host-buffer (e2e) path; 0: the library's default (4 solver waves = 18,944 columns on a B200)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-seq", action="store_true", help="skip the kernel-by-kernel (reference call sequence) leg")
    ap.add_argument("--no-extras", action="store_true", help="skip express / host-pointer / other-config legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer (e2e) leg (profiling runs only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
