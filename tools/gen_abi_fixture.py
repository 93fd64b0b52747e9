#!/usr/bin/env python
"""Extract the extern-mode kernel ABI from the reference's interface-only Fortran files.

Reads  /root/reference/{rte,rrtmgp}/kernels/api/*.F90  (RTE_KERNEL_MODE=extern contract,
rte/kernels/CMakeLists.txt:3-13, rrtmgp/kernels/CMakeLists.txt:3-9) and writes
tests/golden/abi_signatures.json:  { c_symbol: [ [argname, ctype, is_array, intent], ... ] }.

The fixture is committed; tests/test_abi.py compares include/*.h against it, so the check
also runs where /root/reference does not exist (the GPU box).  Re-run this script only when the
reference changes.  It parses declarations itself (it does not use cbind_generator.py, whose
regex drops multi-line declarations).
"""
import glob
import json
import os
import re
import sys

REF = os.environ.get("RTE_RRTMGP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "abi_signatures.json")

TYPE_MAP = {"integer": "int", "real": "Float", "logical": "Bool"}


def join_continuations(src):
    """Strip comments and join '&' continuation lines."""
    out, cur = [], ""
    for raw in src.splitlines():
        line = raw
        # strip comments (no string literals with '!' in these files apart from bind names)
        if "!" in line:
            line = line[: line.index("!")]
        line = line.rstrip()
        if not line.strip():
            continue
        if line.lstrip().startswith("&"):
            line = line.lstrip()[1:]
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        out.append(cur + line)
        cur = ""
    if cur:
        out.append(cur)
    return out


def parse_file(path):
    lines = join_continuations(open(path).read())
    sigs = {}
    i = 0
    sub_re = re.compile(r"^\s*(?:pure\s+)?subroutine\s+(\w+)\s*\((.*?)\)\s*bind\s*\(\s*C\s*,\s*name\s*=\s*\"(\w+)\"\s*\)", re.I)
    while i < len(lines):
        m = sub_re.match(lines[i])
        if not m:
            i += 1
            continue
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        cname = m.group(3)
        decl = {}
        i += 1
        while i < len(lines) and not re.match(r"^\s*end\s+subroutine", lines[i], re.I):
            ln = lines[i]
            dm = re.match(r"^\s*(integer|real|logical)\s*(\([^)]*\))?\s*(.*?)::\s*(.*)$", ln, re.I)
            if dm:
                base = dm.group(1).lower()
                attrs = dm.group(3)
                names = dm.group(4)
                is_arr = bool(re.search(r"dimension\s*\(", attrs, re.I))
                im = re.search(r"intent\s*\(\s*(\w+)\s*\)", attrs, re.I)
                intent = im.group(1).lower() if im else "unspecified"
                # split names at top-level commas (names may carry (dims))
                depth, tok, toks = 0, "", []
                for ch in names:
                    if ch == "(":
                        depth += 1
                    if ch == ")":
                        depth -= 1
                    if ch == "," and depth == 0:
                        toks.append(tok)
                        tok = ""
                    else:
                        tok += ch
                toks.append(tok)
                for t in toks:
                    t = t.strip()
                    if not t:
                        continue
                    nm = re.match(r"(\w+)\s*(\(.*\))?", t)
                    decl[nm.group(1).lower()] = (TYPE_MAP[base], is_arr or bool(nm.group(2)), intent)
            i += 1
        sig = []
        for a in args:
            t = decl.get(a.lower())
            if t is None:
                raise SystemExit(f"{path}: {cname}: no declaration for argument {a}")
            sig.append([a, t[0], t[1], t[2]])
        sigs[cname] = sig
    return sigs


def main():
    allsigs = {}
    for sub in ("rte", "rrtmgp"):
        for f in sorted(glob.glob(os.path.join(REF, sub, "kernels", "api", "*.F90"))):
            s = parse_file(f)
            raw = open(f).read().splitlines()
            for k, v in s.items():
                line = next(i + 1 for i, l in enumerate(raw) if re.search(r'name\s*=\s*"%s"' % k, l))
                allsigs[k] = {"file": os.path.relpath(f, REF), "line": line, "args": v}
    with open(OUT, "w") as fh:
        json.dump(allsigs, fh, indent=1, sort_keys=True)
    print(f"wrote {len(allsigs)} signatures to {os.path.normpath(OUT)}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
