#!/usr/bin/env python3
"""Generates bindings/rust/scir-gpu/src/ffi.rs -- the `extern "C"` block a replacement `scir-gpu` crate binds --
from include/scir_b200.h, one declaration per exported symbol, constants and the plan struct included.

    python tools/gen_rust_ffi.py          # rewrites ffi.rs
    python tools/gen_rust_ffi.py --check  # exit 1 if ffi.rs is stale

tests/test_rust_bindings.py parses the header and ffi.rs INDEPENDENTLY of this script and compares name, arity and the
C type of every argument, so a hand edit of either side that breaks the ABI is caught too.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scir_b200.h")
OUT = os.path.join(ROOT, "bindings", "rust", "scir-gpu", "src", "ffi.rs")

SCALARS = {"int": "c_int", "int64_t": "i64", "uint64_t": "u64", "size_t": "usize", "float": "f32", "double": "f64",
           "char": "c_char", "void": "c_void", "scir_b200_ctx": "ScirB200Ctx", "scir_b200_mg": "ScirB200Mg",
           "scir_b200_resample_plan": "ScirB200ResamplePlan"}


def strip_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def rust_type(ctype):
    """'const float *const *' -> '*const *const f32' (pointer levels read right to left)."""
    t = ctype.strip()
    toks = re.findall(r"\*|const|[A-Za-z_]\w*", t)
    base, base_const, ptrs = None, False, []          # ptrs: list of const-ness of each '*' (pointee constness resolved below)
    i = 0
    while i < len(toks) and toks[i] != "*":
        if toks[i] == "const":
            base_const = True
        else:
            base = toks[i]
        i += 1
    levels = []                                        # const qualifier that FOLLOWS each '*': applies to that pointer itself
    while i < len(toks):
        assert toks[i] == "*", ctype
        i += 1
        c = False
        if i < len(toks) and toks[i] == "const":
            c = True
            i += 1
        levels.append(c)
    r = SCALARS[base]
    pointee_const = base_const
    for self_const in levels:
        r = ("*const " if pointee_const else "*mut ") + r
        pointee_const = self_const
    return r


def parse_header(src):
    src = strip_comments(src)
    protos = []
    for m in re.finditer(r"SCIR_B200_API\s+([^;(]*?)\b(scir_b200_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"^(.*?)(\w+)$", a)
                params.append((mm.group(2), mm.group(1).strip()))
        protos.append((name, ret, params))
    consts = re.findall(r"#define\s+(SCIR_B200_[A-Z_0-9]+)\s+(-?\d+)", src)
    enums = []
    for body in re.findall(r"enum\s*\{(.*?)\}", src, flags=re.S):
        for nm, val in re.findall(r"(SCIR_B200_[A-Z_0-9]+)\s*=\s*(-?\d+)", body):
            enums.append((nm, val))
    return protos, consts, enums


def generate():
    protos, consts, enums = parse_header(open(HEADER).read())
    out = ["//! Raw FFI of libscir_b200.so -- GENERATED from include/scir_b200.h by tools/gen_rust_ffi.py; do not edit.",
           "//! Replaces the hand-declared `extern \"C\"` block over libcuda in the reference crate",
           "//! (crates/scir-gpu/src/lib.rs:549-581): the crate no longer talks to the driver, only to this C ABI.",
           "#![allow(non_camel_case_types, dead_code, missing_docs)]",
           "",
           "use std::os::raw::{c_char, c_int, c_void};",
           "",
           "/// Opaque handle: one device + one stream + scratch (`scir_b200_ctx`).",
           "#[repr(C)]",
           "pub struct ScirB200Ctx {",
           "    _private: [u8; 0],",
           "}",
           "/// Opaque handle: several ctxs, rows sharded across them (`scir_b200_mg`).",
           "#[repr(C)]",
           "pub struct ScirB200Mg {",
           "    _private: [u8; 0],",
           "}",
           "/// Integer plan of `resample_poly` (`scir_b200_resample_plan`).",
           "#[repr(C)]",
           "#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]",
           "pub struct ScirB200ResamplePlan {",
           "    pub up: i64,",
           "    pub down: i64,",
           "    pub n_out: i64,",
           "    pub half_len: i64,",
           "    pub n_pre_pad: i64,",
           "    pub n_post_pad: i64,",
           "    pub n_pre_remove: i64,",
           "    pub len_h_padded: i64,",
           "    pub upfirdn_len: i64,",
           "}",
           ""]
    for nm, val in consts + enums:
        ty = "usize" if nm == "SCIR_B200_MAX_TAPS" else "c_int"
        out.append(f"pub const {nm}: {ty} = {val};")
    out += ["", "#[link(name = \"scir_b200\")]", "extern \"C\" {"]
    for name, ret, params in protos:
        args = ", ".join(f"{p}: {rust_type(t)}" for p, t in params)
        r = rust_type(ret)
        out.append(f"    pub fn {name}({args}) -> {r};")
    out += ["}", ""]
    return "\n".join(out)


if __name__ == "__main__":
    text = generate()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write(text)
    print(f"wrote {OUT}: {text.count('pub fn ')} functions")
