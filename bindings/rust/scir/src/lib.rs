//! scir umbrella crate, GPU part: the re-exports of crates/scir/src/lib.rs:9-14 of SoftOboros/scir, unchanged in
//! name and path, now resolving to the B200 backend.  (The non-GPU re-exports `core`, `fft`, `nd`, `signal` of the
//! reference, lib.rs:4-7, are untouched and omitted here.)
#![deny(missing_docs)]

pub use scir_signal as signal;

#[cfg(feature = "gpu")]
pub mod gpu {
    //! GPU re-exports (enabled with the `gpu` feature).
    pub use scir_gpu::{DType, Device, DeviceArray};
    pub use scir_signal::gpu as signal;
}
