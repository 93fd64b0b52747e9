#!/usr/bin/env python
"""Timeline of the host-buffer streaming driver (RRTMGPB_STREAM_TRACE=1) on the headline workload:
  python tools/stream_trace.py [chunk_cols] [ncol]"""
import os
import sys

os.environ["RRTMGPB_STREAM_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rte_rrtmgp_b200 as pkg  # noqa: E402
from rte_rrtmgp_b200 import synthetic as syn  # noqa: E402
from rte_rrtmgp_b200.frontend import Context  # noqa: E402
from rte_rrtmgp_b200.streaming import HostAllSky  # noqa: E402

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
lib = pkg.lib()
lib.set_stream(torch.cuda.current_stream().cuda_stream)
h = HostAllSky(lib, ncol, 72, syn.make_kdist("lw"), syn.make_kdist("sw"), chunk)
h.step()
Context(lib, "cuda:0").config_checks(False, False)
h.step()
print("--- traced step (checks off) ---", file=sys.stderr)
h.step()
