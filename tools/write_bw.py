#!/usr/bin/env python
"""HBM bandwidth of write-only, read-only and copy streams on this GPU (torch fill_ / sum / copy_ over 8 GiB of float64):
the gas-optics kernels are write-dominated (they emit whole (ncol,nlay,ngpt) planes and read a few KB per cell), so
their roof is what a pure store stream reaches, not the copy figure of MEASURED_PEAKS.json."""
import json
import torch

n = 1 << 30   # 8 GiB of float64
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)


def best(fn, reps=8):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)

ms_w = best(lambda: a.fill_(1.5))
ms_ms = best(lambda: a.zero_())
ms_r = best(lambda: torch.sum(a))
ms_c = best(lambda: b.copy_(a))
gb = n * 8 / 1e9
print(json.dumps({"write_fill_GBps": round(gb / ms_w * 1e3, 1), "write_memset_GBps": round(gb / ms_ms * 1e3, 1),
                  "read_sum_GBps": round(gb / ms_r * 1e3, 1), "copy_GBps_read_plus_write": round(2 * gb / ms_c * 1e3, 1)}))
