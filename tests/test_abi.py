"""The drop-in boundary: include/*.h must declare exactly the reference's extern-mode symbols with the
reference's argument lists (fixture extracted from rte/kernels/api/*.F90 and rrtmgp/kernels/api/*.F90 by
tools/gen_abi_fixture.py), and both shared libraries must export every declared symbol.  No compute calls."""
import ctypes
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIG = json.load(open(os.path.join(ROOT, "tests", "golden", "abi_signatures.json")))
PRODUCT = os.path.join(ROOT, "rte_rrtmgp_b200", "lib", "librte_rrtmgp_b200.so")


def _header_protos(path):
    txt = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(void|int|long long|const char\*|void\*)\s+(\w+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",") if a.strip() and a.strip() != "void"]
        out[m.group(2)] = args
    return out


def test_45_reference_symbols_declared_with_reference_signatures():
    protos = {}
    protos.update(_header_protos(os.path.join(ROOT, "include", "rte_kernels.h")))
    protos.update(_header_protos(os.path.join(ROOT, "include", "rrtmgp_kernels.h")))
    assert len(SIG) == 45
    assert set(protos) == set(SIG)
    for name, ent in SIG.items():
        args = protos[name]
        assert len(args) == len(ent["args"]), name
        for decl, (aname, ctype, _is_arr, intent) in zip(args, ent["args"]):
            # every argument by reference; const iff intent(in); same name and C type
            m = re.fullmatch(r"(const\s+)?(\w+)\s*\*\s*(\w+)", decl)
            assert m, (name, decl)
            assert m.group(2) == ctype and m.group(3) == aname, (name, decl)
            assert bool(m.group(1)) == (intent == "in"), (name, decl)


def test_python_abi_table_matches_fixture():
    from rte_rrtmgp_b200._abi_table import ABI

    assert set(ABI) == set(SIG)
    for name, ent in SIG.items():
        assert [tuple(a) for a in ent["args"]] == [tuple(a) for a in ABI[name]]


def _exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if " T " in ln}


def _declared_everywhere():
    names = set(SIG)
    for h in ("rrtmgp_b200_ext.h", "rrtmgp_b200_frontend.h"):
        names |= set(_header_protos(os.path.join(ROOT, "include", h)))
    return names


def test_oracle_exports_every_declared_symbol(oracle_lib):
    missing = _declared_everywhere() - _exported(oracle_lib.path)
    assert not missing, sorted(missing)
    assert oracle_lib.backend == "cpu-oracle"


@pytest.mark.skipif(not os.path.exists(PRODUCT), reason="product library not built (run __graft_entry__.build())")
def test_product_library_loads_and_exports_every_declared_symbol():
    missing = _declared_everywhere() - _exported(PRODUCT)
    assert not missing, sorted(missing)
    lib = ctypes.CDLL(PRODUCT, mode=ctypes.RTLD_LOCAL)  # loading needs libcudart, not a GPU
    lib.rrtmgpb_backend_name.restype = ctypes.c_char_p
    assert lib.rrtmgpb_backend_name() == b"cuda-sm_100a"


def test_product_never_links_or_loads_the_oracle():
    """The oracle is test infrastructure: nothing under rte_rrtmgp_b200/ may reference it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "rte_rrtmgp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M) or "liboracle" in txt:
                    bad.append(os.path.join(d, f))
    assert not bad, bad
    if os.path.exists(PRODUCT):
        deps = subprocess.run(["ldd", PRODUCT], capture_output=True, text=True).stdout
        assert "oracle" not in deps
