#!/usr/bin/env python3
"""Pin oracle/bsdf.h against a checkout of nvpro_core2.

The BSDF / RNG / light arithmetic the reference path tracer calls lives in nvpro_core2/nvshaders/*.h.slang, which the reference
fetches at configure time (cmake/FindNvproCore2.cmake:28,85, branch `main`, unpinned) and which is NOT under /root/reference —
so oracle/bsdf.h restates it from the published algorithms and says "PARITY UNPINNED".  On a machine where nvpro_core2 IS
reachable, run

    python scripts/pin_against_nvshaders.py --nvpro-core2 /path/to/nvpro_core2 [--out tests/golden/nvshaders_pin.json]

It records the checkout's commit and, for every function of oracle/bsdf.h, (1) whether a function of that name exists in
nvshaders/*.h.slang and in which file, (2) the numeric literals of the two bodies (the constants of the models: 0.04, 1e-4f
thresholds, polynomial coefficients, hash primes ...) and which appear on one side only, (3) the sequence of functions each body
calls.  A clean report (every function found, no one-sided literal, same callee sequence) is what turns "unpinned" into "pinned
at <commit>"; anything else lists what to look at.  The script reads the oracle only as TEXT; nothing in the product imports it.

It cannot run in the build container or on the GPU box (no network, no nvpro_core2): its absence of output there is expected.
"""
import argparse
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

# a definition: optional qualifiers / attributes (static, inline, __device__, PT_HD, [ForceInline] ...), a type, a name, (...) {
FUNC_RE = re.compile(r"^[ \t]*(?:(?:\[\w+\]|[A-Za-z_]\w*)\s+)*[A-Za-z_][\w:<>]*[&*]?\s+([A-Za-z_]\w*)\s*\(([^;{}]*)\)\s*\{", re.M)
NUM_RE = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+|0x[0-9a-fA-F]+|\d{4,})[fFuU]?(?![\w.])")
CALL_RE = re.compile(r"\b([A-Za-z_]\w*)\s*\(")
KEYWORDS = {"if", "for", "while", "return", "switch", "sizeof", "float", "float2", "float3", "float4", "int", "uint", "uint32_t", "f3", "f2",
            "f4", "bool", "static_cast", "max", "min", "fmaxf", "fminf", "sqrtf", "sqrt", "abs", "fabsf", "dot", "cross", "normalize",
            "length", "mix", "lerp", "clamp", "saturate", "sinf", "cosf", "sin", "cos", "powf", "pow", "expf", "exp", "logf", "log",
            "atan2f", "atan2", "acosf", "acos", "floorf", "floor", "any", "all", "asfloat", "asuint", "make_float3"}


def bodies(text):
    """name -> body text (brace matched) for every function definition in `text`."""
    out = {}
    for m in FUNC_RE.finditer(text):
        depth, i = 1, m.end()
        while i < len(text) and depth:
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        if m.group(1) not in KEYWORDS and m.group(1) not in ("else",):
            out.setdefault(m.group(1), text[m.end():i - 1])
    return out


def strip_comments(t):
    return re.sub(r"//[^\n]*|/\*.*?\*/", "", t, flags=re.S)


def literals(body):
    vals = set()
    for tok in NUM_RE.findall(body):
        try:
            vals.add(float(int(tok, 16)) if tok.lower().startswith("0x") else float(tok))
        except ValueError:
            pass
    return vals


def callees(body, known):
    return [c for c in CALL_RE.findall(body) if c in known and c not in KEYWORDS]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--nvpro-core2", required=True, type=Path)
    ap.add_argument("--out", type=Path, default=None)
    a = ap.parse_args()
    shaders = a.nvpro_core2 / "nvshaders"
    if not shaders.is_dir():
        sys.exit(f"{shaders} not found: pass the root of an nvpro_core2 checkout")
    try:
        commit = subprocess.check_output(["git", "-C", str(a.nvpro_core2), "rev-parse", "HEAD"], text=True).strip()
    except Exception:
        commit = "unknown (not a git checkout)"

    ref = {}
    for f in sorted(shaders.glob("*.slang")) + sorted(shaders.glob("*.h")):
        for name, body in bodies(strip_comments(f.read_text(errors="replace"))).items():
            ref.setdefault(name, (f.name, body))
    ours = bodies(strip_comments((ROOT / "oracle" / "bsdf.h").read_text()))
    known = set(ours) | set(ref)

    report, clean = {"nvpro_core2_commit": commit, "functions": {}}, True
    for name, body in ours.items():
        if name not in ref:
            # MDL helpers are prefixed differently in places (mx_/hvd_/...): try a suffix match before giving up
            norm = lambda x: x.replace("_", "").lower()
            cand = [n for n in ref if norm(n) == norm(name)] or [n for n in ref if norm(n).endswith(norm(name)) or norm(name).endswith(norm(n))]
            if len(cand) != 1:
                report["functions"][name] = {"found": False, "candidates": cand}
                clean = False
                continue
            rname = cand[0]
        else:
            rname = name
        fname, rbody = ref[rname]
        lo, lr = literals(body), literals(rbody)
        norm2 = lambda x: x.replace("_", "").lower()
        co, cr = [norm2(c) for c in callees(body, known)], [norm2(c) for c in callees(rbody, known)]
        entry = {"found": True, "file": fname, "reference_name": rname, "only_in_oracle": sorted(lo - lr), "only_in_reference": sorted(lr - lo),
                 "callees_match": co == cr}
        if not entry["callees_match"]:
            entry["callees_oracle"], entry["callees_reference"] = co, cr
        if entry["only_in_oracle"] or entry["only_in_reference"] or not entry["callees_match"]:
            clean = False
        report["functions"][name] = entry
    report["clean"] = clean
    txt = json.dumps(report, indent=1)
    if a.out:
        a.out.write_text(txt + "\n")
    print(txt)
    print(("PINNED at " + commit) if clean else "NOT CLEAN: see the entries above; oracle/bsdf.h stays 'parity unpinned'", file=sys.stderr)
    return 0 if clean else 1


if __name__ == "__main__":
    sys.exit(main())
