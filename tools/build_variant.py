#!/usr/bin/env python
"""Experiment helper: build a VARIANT of the product library with extra -D flags into
rte_rrtmgp_b200/lib/variants/<name>.so (travels to the GPU box; git-ignored like every .so).
Select it at run time with RRTMGPB_LIB=<path> (rte_rrtmgp_b200.lib() honours it for A/B measurements only).

  python tools/build_variant.py v_nohomog -DRB_ADD_HOMOG=0 [--only abi/solvers_abi.cu]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    name = sys.argv[1]
    flags = [a for a in sys.argv[2:] if a.startswith("-D") or a.startswith("-maxrregcount")]
    only = [sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--only"]
    vobj = os.path.join(ROOT, "build", "vobj", name)
    os.makedirs(vobj, exist_ok=True)
    out = os.path.join(ge.PKG, "lib", "variants", name + ".so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    objs, jobs = [], []
    for s in ge._sources():
        rel = os.path.relpath(s, ge.CSRC)
        base = rel.replace(os.sep, "_") + ".o"
        if only and rel not in only:
            objs.append(os.path.join(ge.OBJ, base))  # the main build's object
            continue
        o = os.path.join(vobj, base)
        objs.append(o)
        jobs.append(["/usr/local/cuda/bin/nvcc"] + ge.NVCC_FLAGS + flags + (["-x", "cu"] if s.endswith(".cpp") else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stdout + r.stderr)

    with ThreadPoolExecutor(8) as ex:
        list(ex.map(run, jobs))
    run(["/usr/local/cuda/bin/nvcc", "-shared", "-arch=sm_100a", "-o", out] + objs + ["-Xlinker", "-Bsymbolic", "-lcudart"])
    print(out)


if __name__ == "__main__":
    main()
