//! Builds libscir_b200.so from the CUDA sources with nvcc for sm_100a ONLY and links it.
//!
//! The reference has no build step for its GPU code: a PTX string is JIT-compiled at run time by the driver
//! (crates/scir-gpu/src/lib.rs:814-824).  Here the kernels are compiled ahead of time -- the same nvcc lines as
//! scir_b200/csrc/Makefile -- so a build without nvcc fails loudly instead of producing a crate that silently
//! runs on the CPU.  Set SCIR_B200_LIB_DIR to link a prebuilt library instead.
use std::env;
use std::path::PathBuf;
use std::process::Command;

const SOURCES: &[&str] = &[
    "api.cu", "fir_direct.cu", "fir_direct_rev.cu", "upfirdn.cu", "upfirdn_poly.cu", "fir_toeplitz.cu", "fir_os.cu",
    "fir_f64.cu", "f64_routes.cu", "elementwise.cu", "mg.cu", "microbench.cu",
];

fn main() {
    if env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return; // CPU-only build: Device::Cuda does not exist (lib.rs:28-30), nothing to link
    }
    println!("cargo:rerun-if-env-changed=SCIR_B200_LIB_DIR");
    if let Some(dir) = env::var_os("SCIR_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", PathBuf::from(dir).display());
        println!("cargo:rustc-link-lib=dylib=scir_b200");
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    // <repo>/bindings/rust/scir-gpu -> <repo>/scir_b200/csrc
    let csrc = manifest.join("../../../scir_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let mut objs = Vec::new();
    for src in SOURCES {
        let path = csrc.join(src);
        println!("cargo:rerun-if-changed={}", path.display());
        let obj = out.join(format!("{}.o", src));
        let status = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a"])
            .args(["-Xcompiler", "-fPIC,-fvisibility=hidden", "-c"])
            .arg(&path)
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: the `cuda` feature of scir-gpu needs the CUDA 12.9+ toolkit (sm_100a)");
        assert!(status.success(), "nvcc failed on {}", path.display());
        objs.push(obj);
    }
    let lib = out.join("libscir_b200.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(&lib)
        .args(&objs)
        .args(["-cudart", "static", "-lpthread"])
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "linking libscir_b200.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=scir_b200");
}
