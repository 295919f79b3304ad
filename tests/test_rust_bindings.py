"""The Rust side of the drop-in boundary, checked without a Rust toolchain (none in the image): every symbol
include/scir_b200.h declares is declared in bindings/rust/scir-gpu/src/ffi.rs with the same name, arity and C types,
the constants agree, the safe wrappers keep the reference's public signatures
(crates/scir-gpu/src/lib.rs:515,1036-1039; crates/scir-signal/src/lib.rs:372; crates/scir/src/lib.rs:9-14), and
build.rs compiles the same sources as the Makefile.  The parsers below are independent of tools/gen_rust_ffi.py."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "bindings", "rust")

# canonical C type -> Rust type
BASE = {"int": "c_int", "int64_t": "i64", "uint64_t": "u64", "size_t": "usize", "float": "f32", "double": "f64",
        "char": "c_char", "void": "c_void", "scir_b200_ctx": "ScirB200Ctx", "scir_b200_mg": "ScirB200Mg",
        "scir_b200_resample_plan": "ScirB200ResamplePlan"}


def c_to_rust(ctype):
    """Converts by peeling pointers from the right: T *const * == pointer to (const pointer to T)."""
    t = " ".join(ctype.replace("*", " * ").split())
    if t.endswith("*") or t.endswith("* const"):
        inner = t[: t.rindex("*")].strip()
        # constness of the POINTEE: a trailing `const` of the inner type (T *const) or a leading const of a non-pointer base
        if "*" in inner:
            pointee_const = inner.endswith("const")
            if pointee_const:
                inner = inner[: -len("const")].strip()
        else:
            toks = inner.split()
            pointee_const = "const" in toks
            inner = " ".join(x for x in toks if x != "const")
        return ("*const " if pointee_const else "*mut ") + c_to_rust(inner)
    assert t in BASE, ctype
    return BASE[t]


def header_protos():
    src = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "scir_b200.h")).read(), flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"SCIR_B200_API\s+([^;(]*?)\b(scir_b200_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = " ".join(args.split())
        params = []
        if args != "void":
            for a in args.split(","):
                m = re.match(r"^(.*?)(\w+)$", a.strip())
                params.append((m.group(2), c_to_rust(m.group(1))))
        out[name] = (c_to_rust(ret.strip()), params)
    consts = dict(re.findall(r"#define\s+(SCIR_B200_[A-Z_0-9]+)\s+(-?\d+)", src))
    for body in re.findall(r"enum\s*\{(.*?)\}", src, flags=re.S):
        consts.update(dict(re.findall(r"(SCIR_B200_[A-Z_0-9]+)\s*=\s*(-?\d+)", body)))
    consts.pop("SCIR_B200_H", None)
    return out, consts


def rust_protos():
    src = open(os.path.join(RUST, "scir-gpu", "src", "ffi.rs")).read()
    block = src[src.index('extern "C" {'):]
    out = {}
    for name, args, ret in re.findall(r"pub fn (\w+)\((.*?)\)\s*->\s*([^;]+);", block, flags=re.S):
        params = []
        for a in [p for p in args.split(",") if p.strip()]:
            pn, pt = a.split(":", 1)
            params.append((pn.strip(), " ".join(pt.split())))
        out[name] = (" ".join(ret.split()), params)
    consts = dict(re.findall(r"pub const (SCIR_B200_\w+): \w+ = (-?\d+);", src))
    return out, consts


def test_every_header_symbol_is_bound_with_identical_signature():
    h, hc = header_protos()
    r, rc = rust_protos()
    assert len(h) >= 50
    assert sorted(h) == sorted(r), set(h) ^ set(r)
    for name in h:
        assert h[name][0] == r[name][0], (name, "return type", h[name][0], r[name][0])
        assert len(h[name][1]) == len(r[name][1]), (name, "arity")
        for (hn, ht), (rn, rt) in zip(h[name][1], r[name][1]):
            assert ht == rt, (name, hn, ht, rt)
            assert hn == rn, (name, "parameter name", hn, rn)
    assert hc == rc, set(hc.items()) ^ set(rc.items())


def test_ffi_matches_exported_symbols():
    from scir_b200 import _lib as L
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (scir_b200_\w+)", out)))
    assert exported == sorted(rust_protos()[0])


def test_generator_output_is_committed():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"])
    assert res.returncode == 0, "bindings/rust/scir-gpu/src/ffi.rs is stale: run python tools/gen_rust_ffi.py"


def test_plan_struct_layout_matches():
    src = open(os.path.join(ROOT, "include", "scir_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} scir_b200_resample_plan;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    c_fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            assert decl.startswith("int64_t")
            c_fields += [f.strip() for f in decl[len("int64_t"):].split(",")]
    rs = open(os.path.join(RUST, "scir-gpu", "src", "ffi.rs")).read()
    rbody = re.search(r"pub struct ScirB200ResamplePlan \{(.*?)\}", rs, flags=re.S).group(1)
    r_fields = re.findall(r"pub (\w+): i64", rbody)
    assert c_fields == r_fields
    assert "#[repr(C)]" in rs[: rs.index("pub struct ScirB200ResamplePlan")].rsplit("///", 1)[-1] or "#[repr(C)]" in rs


def test_safe_wrappers_keep_the_reference_signatures():
    lib = " ".join(open(os.path.join(RUST, "scir-gpu", "src", "lib.rs")).read().split())
    # crates/scir-gpu/src/lib.rs:515
    assert "pub fn fir1d_batched_f32_auto(x: &Array2<f32>, taps: &Array1<f32>, device: Device) -> Array2<f32>" in lib
    # crates/scir-gpu/src/lib.rs:1036-1039
    assert "pub fn fir1d_batched_f32_cuda(x: &Array2<f32>, taps: &Array1<f32>) -> Result<Array2<f32>, GpuError>" in lib
    # crates/scir-gpu/src/lib.rs:1134
    assert "pub fn fir1d_batched_f32(x: &Array2<f32>, taps: &Array1<f32>) -> Array2<f32>" in lib
    for item in ("pub enum DType", "pub enum Device", "pub enum GpuError", "pub struct DeviceArray<T>",
                 "pub fn from_cpu_slice(shape: &[usize], dtype: DType, data: &[T]) -> Self", "pub fn to_cpu_vec(&self)",
                 "pub fn to_device(&mut self, device: Device) -> Result<(), GpuError>", "BackendUnavailable(String)",
                 "ShapeMismatch", "pub fn add_scalar_auto(&self, alpha: f32) -> Self",
                 "pub fn add_auto(&self, other: &Self) -> Result<Self, GpuError>", "pub fn mul_scalar_auto(&self, alpha: f32) -> Self"):
        assert item in lib, item
    # no silent fallback: the Cuda arm of _auto must not name the CPU function
    auto = lib[lib.index("pub fn try_fir1d_batched_f32_auto"):]
    auto = auto[: auto.index("#[cfg(feature = \"cuda\")] mod cuda")]
    assert "Device::Cuda => fir1d_batched_f32_cuda(x, taps)" in auto
    sig = " ".join(open(os.path.join(RUST, "scir-signal", "src", "gpu.rs")).read().split())
    # crates/scir-signal/src/lib.rs:372
    assert "pub fn fir1d_batched_f32(x: &Array2<f32>, taps: &Array1<f32>, device: Device) -> Array2<f32>" in sig
    umb = " ".join(open(os.path.join(RUST, "scir", "src", "lib.rs")).read().split())
    # crates/scir/src/lib.rs:9-14
    assert "pub use scir_gpu::{DType, Device, DeviceArray};" in umb and "pub use scir_signal::gpu as signal;" in umb
    cargo = open(os.path.join(RUST, "scir", "Cargo.toml")).read()
    assert re.search(r'^gpu = \["scir-gpu/cuda", "scir-signal/gpu"', cargo, flags=re.M) and 'gpu-all = ["gpu"]' in cargo
    assert re.search(r"^cuda = \[\]", open(os.path.join(RUST, "scir-gpu", "Cargo.toml")).read(), flags=re.M)


def test_wrappers_only_call_declared_ffi_functions():
    declared = set(rust_protos()[0])
    for rel in (("scir-gpu", "src", "lib.rs"), ("scir-signal", "src", "gpu.rs")):
        src = open(os.path.join(RUST, *rel)).read()
        used = set(re.findall(r"ffi::(scir_b200_\w+)", src))
        assert used and used <= declared, used - declared


def test_build_rs_compiles_the_makefile_sources():
    mk = open(os.path.join(ROOT, "scir_b200", "csrc", "Makefile")).read()
    srcs = re.search(r"^SRCS\s*:=\s*(.*)$", mk, flags=re.M).group(1).split()
    b = open(os.path.join(RUST, "scir-gpu", "build.rs")).read()
    listed = re.findall(r'"(\w+\.cu)"', b[b.index("const SOURCES"): b.index("];")])
    assert sorted(listed) == sorted(srcs)
    assert "arch=compute_100a,code=sm_100a" in b
