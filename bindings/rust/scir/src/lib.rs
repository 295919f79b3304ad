//! GPU surface of the `scir` umbrella crate on top of the B200 backend.
//!
//! What stays exactly as in SoftOboros/scir (`crates/scir/src/lib.rs:9-14`, `crates/scir/Cargo.toml:17-19`): with the
//! `gpu` feature, `scir::gpu::{DType, Device, DeviceArray}` and `scir::gpu::signal` exist under these names, so
//!
//! ```ignore
//! use scir::gpu::{signal, Device};
//! let y = signal::fir1d_batched_f32(&x, &taps, Device::Cuda);      // same call, same shapes -- now a B200 kernel
//! ```
//!
//! keeps compiling.  What is new and additive: the error type, the `Result`-returning sibling of the dispatch
//! function (the reference swallowed CUDA errors and fell back to the CPU; this backend never does), the explicit
//! context handle and the in-process multi-GPU front end.  The umbrella's non-GPU re-exports (`core`, `fft`, `nd`,
//! `signal`; reference `lib.rs:4-7`) are untouched by this backend and not repeated here.
#![deny(missing_docs)]

/// `scir::signal` -- the signal crate, as before.
pub use scir_signal as signal;

/// GPU re-exports (feature `gpu`, which turns on `scir-gpu/cuda` and `scir-signal/gpu`).
#[cfg(feature = "gpu")]
pub mod gpu {
    // -- the reference's surface, names unchanged -------------------------------------------------------------------
    pub use scir_gpu::{DType, Device, DeviceArray};
    pub use scir_signal::gpu as signal;

    // -- additive: errors are reported, not swallowed ----------------------------------------------------------------
    pub use scir_gpu::{fir1d_batched_f32_auto, fir1d_batched_f32_cuda, try_fir1d_batched_f32_auto, GpuError};

    // -- additive: long-lived device context and row sharding over the GPUs of one box -------------------------------
    pub use scir_gpu::{with_default_context, Context, MultiGpu};
}
