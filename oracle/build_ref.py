#!/usr/bin/env python3
"""Build oracle/_ref/ from the REFERENCE's own sources, where they lie.

TEST INFRASTRUCTURE ONLY (never imported by the product path).

What is built (all outputs go to oracle/_ref/, which is git-ignored but travels to the
GPU box with gpurun snapshots):

  libfr_ref.so     reference field arithmetic: code_producers/src/c_elements/generic/
                   fr.cpp + fr.hpp, rendered for BN254 exactly as
                   code_producers/src/c_elements/c_code_generator.rs:1015-1067 renders
                   them (handlebars placeholders), + oracle/fr_ref_driver.cpp.
  calcwit.o main.o reference runtime code_producers/src/c_elements/common/{calcwit,main}.cpp,
                   compiled unchanged (the `main` symbol of main.o is renamed with objcopy
                   so that our timing harness can own main()).
  fr.o             the rendered fr.cpp as an object for linking circuit binaries.

Not buildable here (stated in DESIGN.md): bn128/fr.asm (no nasm) and the Rust compiler
(no cargo) -- so the circuit bodies linked against this runtime are emitted by
tools/circuitgen in the shape of the reference's WriteC emitters.

The gmp.h used is oracle/gmp_shim/gmp.h (prototypes only; links to the system
libgmp.so.10).  nlohmann/json.hpp is taken from the cudnn-frontend third-party dir of
this image's site-packages.
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CIRCOM_REFERENCE", "/root/reference")
CEL = os.path.join(REF, "code_producers/src/c_elements")
OUT = os.path.join(HERE, "_ref")
LIBGMP = "/usr/lib/x86_64-linux-gnu/libgmp.so.10"

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _u64_list(x, n64):
    return ["0x%x" % ((x >> (64 * i)) & ((1 << 64) - 1)) for i in range(n64)]


def bn254_context():
    """Same quantities as c_code_generator.rs:1028-1067 computes for the template."""
    p = Q
    pbits = p.bit_length()
    n64 = (pbits + 63) // 64
    nbits = n64 * 64
    half = p // 2
    inv = pow(p, -1, 1 << 64)
    np_ = (1 << 64) - inv
    lbo_mask = ((1 << 64) >> (nbits - pbits)) - 1
    r2 = (1 << (nbits * 2)) % p
    r3 = (1 << (nbits * 3)) % p
    return {
        "cannotOptimize": (p >> ((n64 - 1) * 64)) > ((((1 << 64) - 1) >> 1) - 1),
        "list0n64": list(range(n64)),
        "list0n64_1": list(range(n64 - 1)),
        "list1n64": list(range(1, n64)),
        "n64": n64,
        "fr_n64": n64,
        "qbits": pbits,
        "lboMask": "0x%x" % lbo_mask,
        "fr_np": "0x%x" % np_,
        "fr_q_list": _u64_list(p, n64),
        "fr_r2_list": _u64_list(r2, n64),
        "fr_r3_list": _u64_list(r3, n64),
        "half_list": _u64_list(half, n64),
    }


# --- a small handlebars subset: {{x}}, {{elements x}}, {{inc e}}, {{dec e}}, {{@index}},
# --- {{this}}, {{#if c}}..{{else}}..{{/if}}, {{#each l}}..{{/each}}, {{#if @last}}
_TOK = re.compile(r"\{\{\s*(.*?)\s*\}\}", re.S)


def _eval_expr(expr, ctx, loc):
    expr = expr.strip()
    if expr.startswith("(") and expr.endswith(")"):
        return _eval_expr(expr[1:-1], ctx, loc)
    m = re.match(r"^(inc|dec|elements)\s+(.*)$", expr, re.S)
    if m:
        v = _eval_expr(m.group(2), ctx, loc)
        if m.group(1) == "inc":
            return int(v) + 1
        if m.group(1) == "dec":
            return int(v) - 1
        return ",".join(v)
    if expr == "@index":
        return loc["index"]
    if expr == "@last":
        return loc["last"]
    if expr == "this":
        return loc["this"]
    return ctx[expr]


def _parse(tokens, pos, stop):
    """tokens: list of ('text', s) | ('tag', expr).  Returns (nodes, pos)."""
    nodes = []
    while pos < len(tokens):
        kind, val = tokens[pos]
        if kind == "text":
            nodes.append(("text", val))
            pos += 1
            continue
        if val in stop:
            return nodes, pos
        if val.startswith("#if"):
            cond = val[3:].strip()
            then, pos = _parse(tokens, pos + 1, {"else", "/if"})
            other = []
            if tokens[pos][1] == "else":
                other, pos = _parse(tokens, pos + 1, {"/if"})
            nodes.append(("if", cond, then, other))
            pos += 1
        elif val.startswith("#each"):
            lst = val[5:].strip()
            body, pos = _parse(tokens, pos + 1, {"/each"})
            nodes.append(("each", lst, body))
            pos += 1
        else:
            nodes.append(("expr", val))
            pos += 1
    return nodes, pos


def _emit(nodes, ctx, loc, out):
    for n in nodes:
        if n[0] == "text":
            out.append(n[1])
        elif n[0] == "expr":
            out.append(str(_eval_expr(n[1], ctx, loc)))
        elif n[0] == "if":
            branch = n[2] if _eval_expr(n[1], ctx, loc) else n[3]
            _emit(branch, ctx, loc, out)
        elif n[0] == "each":
            lst = _eval_expr(n[1], ctx, loc)
            for i, item in enumerate(lst):
                _emit(n[2], ctx, {"index": i, "this": item, "last": i == len(lst) - 1}, out)


def render(template, ctx):
    tokens = []
    last = 0
    for m in _TOK.finditer(template):
        if m.start() > last:
            tokens.append(("text", template[last:m.start()]))
        tokens.append(("tag", m.group(1).strip()))
        last = m.end()
    tokens.append(("text", template[last:]))
    nodes, _ = _parse(tokens, 0, set())
    out = []
    _emit(nodes, ctx, {}, out)
    res = "".join(out)
    assert "{{" not in res, "unrendered placeholder left"
    return res


def sh(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def find_json_hpp():
    site = sysconfig.get_paths()["purelib"]
    cand = os.path.join(site, "include/cudnn_frontend/thirdparty/nlohmann/json.hpp")
    if os.path.exists(cand):
        return cand
    for root, _d, files in os.walk(site):
        if "json.hpp" in files and root.endswith("nlohmann"):
            return os.path.join(root, "json.hpp")
    raise FileNotFoundError("nlohmann/json.hpp")


def build():
    if not os.path.isdir(CEL):
        print("reference tree not present (%s); keeping prebuilt oracle/_ref" % CEL)
        return False
    src = os.path.join(OUT, "src")
    inc = os.path.join(OUT, "inc", "nlohmann")
    # up to date: every product newer than every input (reference sources where they lie, this recipe, the driver)
    products = [os.path.join(OUT, f) for f in ("fr.o", "calcwit.o", "main.o", "libfr_ref.so")] + \
               [os.path.join(src, f) for f in ("fr.cpp", "fr.hpp", "circom.hpp", "calcwit.hpp")]
    inputs = [os.path.join(CEL, "generic", "fr.cpp"), os.path.join(CEL, "generic", "fr.hpp"),
              os.path.join(CEL, "common", "calcwit.cpp"), os.path.join(CEL, "common", "main.cpp"),
              os.path.join(CEL, "common", "calcwit.hpp"), os.path.join(CEL, "common", "circom.hpp"),
              os.path.abspath(__file__), os.path.join(HERE, "fr_ref_driver.cpp"), os.path.join(HERE, "gmp_shim", "gmp.h")]
    if all(os.path.exists(p) for p in products) and \
            min(os.path.getmtime(p) for p in products) > max(os.path.getmtime(p) for p in inputs):
        print("oracle/_ref runtime is up to date")
        return True
    os.makedirs(src, exist_ok=True)
    os.makedirs(inc, exist_ok=True)
    ctx = bn254_context()
    for name in ("fr.cpp", "fr.hpp"):
        with open(os.path.join(CEL, "generic", name)) as f:
            text = render(f.read(), ctx)
        with open(os.path.join(src, name), "w") as f:
            f.write(text)
    shutil.copy(find_json_hpp(), os.path.join(inc, "json.hpp"))
    common = os.path.join(CEL, "common")
    flags = ["-std=c++11", "-O3", "-fPIC", "-Wno-address-of-packed-member", "-w",
             "-I", src, "-I", os.path.join(HERE, "gmp_shim"), "-I", common,
             "-I", os.path.join(OUT, "inc")]
    sh(["g++", *flags, "-c", os.path.join(src, "fr.cpp"), "-o", os.path.join(OUT, "fr.o")])
    sh(["g++", *flags, "-c", os.path.join(common, "calcwit.cpp"), "-o", os.path.join(OUT, "calcwit.o")])
    # main.cpp is compiled unchanged; its `main` symbol is then renamed so that oracle/ref_harness.cpp can own main()
    # (a -Dmain=... macro would also rewrite identifiers inside the standard headers and breaks iostream).
    sh(["g++", *flags, "-include", "cstring", "-include", "cassert",
        "-c", os.path.join(common, "main.cpp"), "-o", os.path.join(OUT, "main_plain.o")])
    sh(["objcopy", "--redefine-sym", "main=circom_reference_main", os.path.join(OUT, "main_plain.o"),
        os.path.join(OUT, "main.o")])
    sh(["g++", *flags, "-shared", os.path.join(HERE, "fr_ref_driver.cpp"), os.path.join(OUT, "fr.o"),
        LIBGMP, "-o", os.path.join(OUT, "libfr_ref.so")])
    # headers needed later to compile emitted circuit bodies against the runtime
    for name in ("circom.hpp", "calcwit.hpp"):
        shutil.copy(os.path.join(common, name), os.path.join(src, name))
    return True


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok or os.path.exists(os.path.join(OUT, "libfr_ref.so")) else 1)
