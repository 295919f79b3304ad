//! GPU Foundations on B200: device arrays and batched FIR through the `scir_b200` C ABI.
//!
//! Drop-in replacement for crate `scir-gpu` of SoftOboros/scir for ONE path -- batched real-f32 FIR filtering and
//! the `DeviceArray` plumbing around it.  Public names, argument meaning and output shapes follow the reference:
//!
//! | here | reference (`crates/scir-gpu/src/lib.rs`) |
//! |---|---|
//! | [`DType`], [`Device`], [`GpuError`] | `:16-22`, `:25-35`, `:57-74` |
//! | [`DeviceArray`] | `:77-190` (CPU-backed placeholder there; device storage here) |
//! | [`fir1d_batched_f32`] (CPU) | `:1134-1152` |
//! | [`fir1d_batched_f32_cuda`] | `:1036-1113` + PTX `:727-811` |
//! | [`fir1d_batched_f32_auto`] | `:515-531` |
//!
//! Differences, all required by the B200 north star: `Device::Cuda` NEVER falls back to the CPU (the reference
//! swallowed the CUDA error and ran the CPU loop, `:520-523`); `fir1d_batched_f32_auto` keeps its infallible signature,
//! so it panics with the backend's message when the GPU path cannot run, and [`try_fir1d_batched_f32_auto`] is the
//! `Result` sibling.  There is no `wgpu` arm.  `Device::Cpu`, asked for explicitly, runs the crate's own CPU loop as
//! before (`:517-518`).
#![deny(missing_docs)]

use ndarray::{Array1, Array2};
use std::error::Error;
use std::fmt;

/// Raw `extern "C"` declarations of `libscir_b200.so` (generated from `include/scir_b200.h`).
#[cfg(feature = "cuda")]
pub mod ffi;

/// Supported data types for device arrays (`lib.rs:16-22`).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum DType {
    /// 32-bit floating point
    F32,
    /// 64-bit floating point
    F64,
}

/// Execution device selection (`lib.rs:25-35`).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Device {
    /// Host CPU device
    Cpu,
    #[cfg(feature = "cuda")]
    /// NVIDIA B200 through libscir_b200.so (feature `cuda`)
    Cuda,
}

/// GPU-related error types (`lib.rs:57-74`).
#[derive(Debug)]
pub enum GpuError {
    /// Backend is not available on this build or platform (no B200, no driver, launch failure ...).
    BackendUnavailable(String),
    /// Operation failed due to incompatible shapes.
    ShapeMismatch,
}

impl fmt::Display for GpuError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        match self {
            GpuError::BackendUnavailable(name) => write!(f, "backend not available: {name}"),
            GpuError::ShapeMismatch => write!(f, "shape mismatch"),
        }
    }
}

impl Error for GpuError {}

/// Causal FIR over each row of `x` using `taps` (CPU, f32): `y[b,i] = sum_t taps[k-1-t] * x[b,i-t]`, zero initial
/// state, f32 multiply then f32 add, newest sample first (`lib.rs:1134-1152`).  Input shape is `(batch, n)` and the
/// same shape is returned.  This is the `Device::Cpu` arm of [`fir1d_batched_f32_auto`]; the `Device::Cuda` arm never
/// reaches it.
pub fn fir1d_batched_f32(x: &Array2<f32>, taps: &Array1<f32>) -> Array2<f32> {
    let (b, n) = x.dim();
    let k = taps.len();
    let mut y = Array2::<f32>::zeros((b, n));
    for bi in 0..b {
        for i in 0..n {
            let mut acc = 0.0f32;
            let reach = if i + 1 < k { i + 1 } else { k };
            for t in 0..reach {
                acc += taps[k - 1 - t] * x[[bi, i - t]];
            }
            y[[bi, i]] = acc;
        }
    }
    y
}

/// Batched FIR with device selection (`lib.rs:515-531`), WITHOUT the silent fallback: `Device::Cuda` runs on the
/// B200 or panics with the backend's error text (the signature has no `Result`; use
/// [`try_fir1d_batched_f32_auto`] to handle the error).
pub fn fir1d_batched_f32_auto(x: &Array2<f32>, taps: &Array1<f32>, device: Device) -> Array2<f32> {
    match try_fir1d_batched_f32_auto(x, taps, device) {
        Ok(y) => y,
        Err(e) => panic!("fir1d_batched_f32_auto: {e}"),
    }
}

/// `Result` sibling of [`fir1d_batched_f32_auto`].
pub fn try_fir1d_batched_f32_auto(
    x: &Array2<f32>,
    taps: &Array1<f32>,
    device: Device,
) -> Result<Array2<f32>, GpuError> {
    match device {
        Device::Cpu => Ok(fir1d_batched_f32(x, taps)),
        #[cfg(feature = "cuda")]
        Device::Cuda => fir1d_batched_f32_cuda(x, taps),
    }
}

#[cfg(feature = "cuda")]
mod cuda {
    use super::*;
    use crate::ffi;
    use std::cell::RefCell;
    use std::ffi::CStr;
    use std::os::raw::c_void;
    use std::ptr;

    /// Maps an ABI return code to `GpuError`.
    pub(crate) fn check(rc: std::os::raw::c_int) -> Result<(), GpuError> {
        if rc == ffi::SCIR_B200_OK {
            return Ok(());
        }
        if rc == ffi::SCIR_B200_ERR_SHAPE {
            return Err(GpuError::ShapeMismatch);
        }
        let msg = unsafe { CStr::from_ptr(ffi::scir_b200_last_error()) }.to_string_lossy().into_owned();
        Err(GpuError::BackendUnavailable(format!("scir_b200 error {rc}: {msg}")))
    }

    /// A long-lived handle (device + stream + scratch); replaces the per-call `CudaCtx` (`lib.rs:601-622`).
    pub struct Context {
        raw: *mut ffi::ScirB200Ctx,
    }

    impl Context {
        /// Creates a ctx on `device`; `Err(BackendUnavailable)` without a B200.
        pub fn new(device: i32) -> Result<Self, GpuError> {
            let mut raw = ptr::null_mut();
            check(unsafe { ffi::scir_b200_ctx_create(device, &mut raw) })?;
            Ok(Self { raw })
        }
        /// Creates a ctx on the calling thread's current CUDA device.
        pub fn on_current_device() -> Result<Self, GpuError> {
            let mut dev = 0;
            check(unsafe { ffi::scir_b200_current_device(&mut dev) })?;
            Self::new(dev)
        }
        /// The raw handle, for the `scir_b200_*` entry points not wrapped here.
        pub fn as_ptr(&self) -> *mut ffi::ScirB200Ctx {
            self.raw
        }
        /// Waits for the ctx's stream.
        pub fn sync(&self) -> Result<(), GpuError> {
            check(unsafe { ffi::scir_b200_ctx_sync(self.raw) })
        }
    }

    impl Drop for Context {
        fn drop(&mut self) {
            unsafe {
                ffi::scir_b200_ctx_destroy(self.raw);
            }
        }
    }

    thread_local! {
        static DEFAULT_CTX: RefCell<Option<Context>> = RefCell::new(None);
    }

    /// Runs `f` with this thread's default ctx (created on first use, on the thread's current device).
    pub fn with_default_context<R>(f: impl FnOnce(&Context) -> Result<R, GpuError>) -> Result<R, GpuError> {
        DEFAULT_CTX.with(|slot| {
            let mut slot = slot.borrow_mut();
            if slot.is_none() {
                *slot = Some(Context::on_current_device()?);
            }
            f(slot.as_ref().unwrap())
        })
    }

    /// CUDA 1D batched FIR on the B200 (`lib.rs:1036-1113`): rows are streamed through the device in blocks, H2D /
    /// kernel / D2H overlapped; pageable `Array2` storage goes through the ctx's pinned ring.
    ///
    /// # Errors
    /// [`GpuError::BackendUnavailable`] if there is no B200 or a CUDA call fails.  Never falls back to the CPU.
    pub fn fir1d_batched_f32_cuda(x: &Array2<f32>, taps: &Array1<f32>) -> Result<Array2<f32>, GpuError> {
        let (b, n) = x.dim();
        let x_std = x.as_standard_layout(); // the reference deep-copies too (lib.rs:1042); a no-op for C-contiguous input
        let x_host = x_std.as_slice().ok_or(GpuError::ShapeMismatch)?;
        let taps_std = taps.as_standard_layout();
        let taps_host = taps_std.as_slice().ok_or(GpuError::ShapeMismatch)?;
        let mut out_host = vec![0.0f32; b * n];
        let ld = if n > 0 { n as i64 } else { 1 };
        with_default_context(|ctx| {
            check(unsafe {
                ffi::scir_b200_fir1d_batched_f32_host(
                    ctx.as_ptr(),
                    x_host.as_ptr(),
                    ld,
                    taps_host.as_ptr(),
                    taps_host.len() as i64,
                    ffi::SCIR_B200_TAPS_SCIR,
                    out_host.as_mut_ptr(),
                    ld,
                    b as i64,
                    n as i64,
                )
            })
        })?;
        Array2::from_shape_vec((b, n), out_host).map_err(|_| GpuError::ShapeMismatch)
    }

    /// The in-process multi-GPU front end (north_star (d)): rows (channels) of one host array sharded over several
    /// devices, one stream and one host thread per device, no collective (`scir_b200_mg_*`).
    pub struct MultiGpu {
        raw: *mut ffi::ScirB200Mg,
    }

    impl MultiGpu {
        /// One shard per entry of `devices` (a device may be listed more than once).
        pub fn new(devices: &[i32]) -> Result<Self, GpuError> {
            let mut raw = ptr::null_mut();
            check(unsafe { ffi::scir_b200_mg_create(devices.as_ptr(), devices.len() as i32, &mut raw) })?;
            Ok(Self { raw })
        }
        /// All devices of the box.
        pub fn all_devices() -> Result<Self, GpuError> {
            let mut n = 0;
            check(unsafe { ffi::scir_b200_device_count(&mut n) })?;
            let devs: Vec<i32> = (0..n).collect();
            Self::new(&devs)
        }
        /// `fir1d_batched_f32` with the rows sharded over the devices; same values as the single-device call.
        pub fn fir1d_batched_f32(&self, x: &Array2<f32>, taps: &Array1<f32>) -> Result<Array2<f32>, GpuError> {
            let (b, n) = x.dim();
            let x_std = x.as_standard_layout();
            let x_host = x_std.as_slice().ok_or(GpuError::ShapeMismatch)?;
            let taps_std = taps.as_standard_layout();
            let taps_host = taps_std.as_slice().ok_or(GpuError::ShapeMismatch)?;
            let mut out_host = vec![0.0f32; b * n];
            let ld = if n > 0 { n as i64 } else { 1 };
            check(unsafe {
                ffi::scir_b200_mg_fir1d_batched_f32_host(
                    self.raw, x_host.as_ptr(), ld, taps_host.as_ptr(), taps_host.len() as i64, ffi::SCIR_B200_TAPS_SCIR,
                    out_host.as_mut_ptr(), ld, b as i64, n as i64,
                )
            })?;
            Array2::from_shape_vec((b, n), out_host).map_err(|_| GpuError::ShapeMismatch)
        }
    }

    impl Drop for MultiGpu {
        fn drop(&mut self) {
            unsafe {
                ffi::scir_b200_mg_destroy(self.raw);
            }
        }
    }

    /// Device storage of a [`super::DeviceArray<f32>`]: owned by the default ctx of the creating thread.
    pub(crate) struct DeviceBuf {
        pub(crate) ptr: *mut c_void,
        pub(crate) len: usize,
    }

    impl DeviceBuf {
        pub(crate) fn alloc(len: usize) -> Result<Self, GpuError> {
            let mut raw: *mut c_void = ptr::null_mut();
            with_default_context(|ctx| {
                check(unsafe { ffi::scir_b200_malloc(ctx.as_ptr(), len.max(1) * 4, &mut raw) })
            })?;
            Ok(Self { ptr: raw, len })
        }
        pub(crate) fn upload(host: &[f32]) -> Result<Self, GpuError> {
            let buf = Self::alloc(host.len())?;
            with_default_context(|ctx| {
                check(unsafe {
                    ffi::scir_b200_memcpy_h2d(ctx.as_ptr(), buf.ptr, host.as_ptr() as *const c_void, host.len() * 4)
                })
            })?;
            Ok(buf)
        }
        pub(crate) fn download(&self) -> Result<Vec<f32>, GpuError> {
            let mut out = vec![0.0f32; self.len];
            with_default_context(|ctx| {
                check(unsafe {
                    ffi::scir_b200_memcpy_d2h(ctx.as_ptr(), out.as_mut_ptr() as *mut c_void, self.ptr, self.len * 4)
                })
            })?;
            Ok(out)
        }
    }

    impl Drop for DeviceBuf {
        fn drop(&mut self) {
            let p = self.ptr;
            let _ = with_default_context(|ctx| check(unsafe { ffi::scir_b200_free(ctx.as_ptr(), p) }));
        }
    }
}

#[cfg(feature = "cuda")]
pub use cuda::{fir1d_batched_f32_cuda, with_default_context, Context, MultiGpu};

/// A shaped array with dtype that lives on the host or -- after `to_device(Device::Cuda)` -- in B200 memory
/// (`lib.rs:77-190`, where the CUDA arm only re-tagged a host `Vec`).  Device storage is f32 (the FIR path's dtype).
pub struct DeviceArray<T> {
    shape: Vec<usize>,
    dtype: DType,
    device: Device,
    host: Vec<T>,
    #[cfg(feature = "cuda")]
    dev: Option<cuda::DeviceBuf>,
}

impl<T: Copy> DeviceArray<T> {
    /// Create a `DeviceArray` from a CPU slice and explicit shape/dtype (`lib.rs:95-103`).
    pub fn from_cpu_slice(shape: &[usize], dtype: DType, data: &[T]) -> Self {
        assert_eq!(shape.iter().product::<usize>(), data.len());
        Self {
            shape: shape.to_vec(),
            dtype,
            device: Device::Cpu,
            host: data.to_vec(),
            #[cfg(feature = "cuda")]
            dev: None,
        }
    }
    /// Return the logical shape of the array.
    pub fn shape(&self) -> &[usize] {
        &self.shape
    }
    /// Return the element data type.
    pub fn dtype(&self) -> DType {
        self.dtype
    }
    /// Return the current device of this array.
    pub fn device(&self) -> Device {
        self.device
    }
}

impl DeviceArray<f32> {
    /// Copy data back to a CPU-owned `Vec` (`lib.rs:114-116`); a real D2H copy for arrays on `Device::Cuda`.
    pub fn to_cpu_vec(&self) -> Vec<f32> {
        #[cfg(feature = "cuda")]
        if let Some(buf) = &self.dev {
            return buf.download().unwrap_or_else(|e| panic!("DeviceArray::to_cpu_vec: {e}"));
        }
        self.host.clone()
    }

    /// Move the array to a device (`lib.rs:158-189`): a real upload / download here.
    pub fn to_device(&mut self, device: Device) -> Result<(), GpuError> {
        if device == self.device {
            return Ok(());
        }
        match device {
            Device::Cpu => {
                #[cfg(feature = "cuda")]
                if let Some(buf) = self.dev.take() {
                    self.host = buf.download()?;
                }
                self.device = Device::Cpu;
                Ok(())
            }
            #[cfg(feature = "cuda")]
            Device::Cuda => {
                self.dev = Some(cuda::DeviceBuf::upload(&self.host)?);
                self.device = Device::Cuda;
                Ok(())
            }
        }
    }

    fn host_like(&self, host: Vec<f32>) -> Self {
        Self {
            shape: self.shape.clone(),
            dtype: self.dtype,
            device: Device::Cpu,
            host,
            #[cfg(feature = "cuda")]
            dev: None,
        }
    }

    /// Add a scalar with device dispatch (`lib.rs:268-301`).  The array's device decides; no arm falls back to another.
    pub fn add_scalar_auto(&self, alpha: f32) -> Self {
        #[cfg(feature = "cuda")]
        if let Some(buf) = &self.dev {
            return self.device_unary(buf, alpha, true);
        }
        self.host_like(self.host.iter().map(|v| *v + alpha).collect())
    }

    /// Multiply by a scalar with device dispatch (`lib.rs:355-388`).
    pub fn mul_scalar_auto(&self, alpha: f32) -> Self {
        #[cfg(feature = "cuda")]
        if let Some(buf) = &self.dev {
            return self.device_unary(buf, alpha, false);
        }
        self.host_like(self.host.iter().map(|v| *v * alpha).collect())
    }

    /// Elementwise sum with device dispatch (`lib.rs:303-353`); shapes must match.
    pub fn add_auto(&self, other: &Self) -> Result<Self, GpuError> {
        if self.shape != other.shape {
            return Err(GpuError::ShapeMismatch);
        }
        #[cfg(feature = "cuda")]
        if let (Some(a), Some(b)) = (&self.dev, &other.dev) {
            let out = cuda::DeviceBuf::alloc(a.len)?;
            with_default_context(|ctx| {
                cuda::check(unsafe {
                    ffi::scir_b200_add_f32(ctx.as_ptr(), a.ptr as *const f32, b.ptr as *const f32, out.ptr as *mut f32, a.len as i64)
                })
            })?;
            return Ok(self.device_like(out));
        }
        if self.device != other.device {
            return Err(GpuError::BackendUnavailable("add_auto: operands live on different devices".into()));
        }
        Ok(self.host_like(self.host.iter().zip(other.host.iter()).map(|(a, b)| *a + *b).collect()))
    }

    #[cfg(feature = "cuda")]
    fn device_like(&self, buf: cuda::DeviceBuf) -> Self {
        Self { shape: self.shape.clone(), dtype: self.dtype, device: Device::Cuda, host: Vec::new(), dev: Some(buf) }
    }

    #[cfg(feature = "cuda")]
    fn device_unary(&self, buf: &cuda::DeviceBuf, alpha: f32, add: bool) -> Self {
        let out = cuda::DeviceBuf::alloc(buf.len).unwrap_or_else(|e| panic!("DeviceArray: {e}"));
        let rc = with_default_context(|ctx| {
            cuda::check(unsafe {
                if add {
                    ffi::scir_b200_add_scalar_f32(ctx.as_ptr(), buf.ptr as *const f32, alpha, out.ptr as *mut f32, buf.len as i64)
                } else {
                    ffi::scir_b200_mul_scalar_f32(ctx.as_ptr(), buf.ptr as *const f32, alpha, out.ptr as *mut f32, buf.len as i64)
                }
            })
        });
        if let Err(e) = rc {
            panic!("DeviceArray elementwise op: {e}");
        }
        self.device_like(out)
    }

    /// Device-resident batched FIR on a 2-D array living on `Device::Cuda`: no PCIe traffic when chaining
    /// (SURVEY.md 8(f).1).  `taps` in the reference's order.
    #[cfg(feature = "cuda")]
    pub fn fir1d_batched(&self, taps: &Array1<f32>) -> Result<Self, GpuError> {
        let buf = self.dev.as_ref().ok_or(GpuError::ShapeMismatch)?;
        if self.shape.len() != 2 {
            return Err(GpuError::ShapeMismatch);
        }
        let (b, n) = (self.shape[0], self.shape[1]);
        let taps_std = taps.as_standard_layout();
        let t = taps_std.as_slice().ok_or(GpuError::ShapeMismatch)?;
        let out = cuda::DeviceBuf::alloc(b * n)?;
        let ld = if n > 0 { n as i64 } else { 1 };
        with_default_context(|ctx| {
            cuda::check(unsafe {
                ffi::scir_b200_fir1d_batched_f32(
                    ctx.as_ptr(), buf.ptr as *const f32, ld, t.as_ptr(), t.len() as i64, ffi::SCIR_B200_TAPS_SCIR,
                    out.ptr as *mut f32, ld, b as i64, n as i64,
                )
            })
        })?;
        Ok(self.device_like(out))
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use ndarray::array;

    // the reference's known-answer vector (crates/scir-gpu/src/lib.rs:1251-1260)
    #[test]
    fn fir1d_batched_f32_known_answer() {
        let x: Array2<f32> = array![[1.0, 2.0, 3.0, 4.0], [0.5, 0.0, -0.5, -1.0]];
        let taps: Array1<f32> = array![0.25, 0.5, 0.25];
        let y = fir1d_batched_f32(&x, &taps);
        let want: Array2<f32> = array![[0.25, 1.0, 2.0, 3.0], [0.125, 0.25, 0.0, -0.5]];
        for (a, b) in y.iter().zip(want.iter()) {
            assert!((a - b).abs() <= 1e-7);
        }
    }

    // crates/scir-gpu/src/lib.rs:1300-1323, with the difference that Err is a FAILURE here (the reference's test passed on Err)
    #[cfg(feature = "cuda")]
    #[test]
    fn cuda_fir1d_batched_f32_parity_small() {
        let x: Array2<f32> = array![[1.0, 2.0, 3.0, 4.0], [0.5, 0.0, -0.5, -1.0]];
        let taps: Array1<f32> = array![0.25, 0.5, 0.25];
        let y_cpu = fir1d_batched_f32(&x, &taps);
        let y_gpu = fir1d_batched_f32_cuda(&x, &taps).expect("Device::Cuda must run on a B200 box");
        for (a, b) in y_cpu.iter().zip(y_gpu.iter()) {
            assert!((a - b).abs() <= 1e-5 + 1e-6 * a.abs());
        }
    }
}
