//! `scir_signal::gpu` -- GPU-forwarded FIR routes (feature `gpu`), the replacement for
//! crates/scir-signal/src/lib.rs:365-375 of SoftOboros/scir.  Mount with
//!
//! ```ignore
//! #[cfg(feature = "gpu")]
//! #[path = "gpu.rs"]
//! pub mod gpu;
//! ```
//!
//! `fir1d_batched_f32` keeps the reference's signature (sig/lib.rs:372).  The other functions are the FIR routes
//! north_star (c) maps onto the same kernels -- `lfilter` with `a = [1]`, `upfirdn` / `resample_poly`, FIR `filtfilt`
//! -- with SciPy's semantics (the reference's own `resample_poly` / `filtfilt`, sig/lib.rs:278-362, are f64 and
//! CPU-only; their f64 twins are `resample_poly_f64` / `filtfilt_fir_f64` below).
use ndarray::{Array1, Array2};
use scir_gpu::{fir1d_batched_f32_auto, Device, GpuError};

/// Batched FIR for f32 with device selection (sig/lib.rs:372-374).
pub fn fir1d_batched_f32(x: &Array2<f32>, taps: &Array1<f32>, device: Device) -> Array2<f32> {
    fir1d_batched_f32_auto(x, taps, device)
}

#[cfg(feature = "cuda")]
mod routes {
    use super::*;
    use scir_gpu::ffi;
    use scir_gpu::with_default_context;
    use std::ffi::CStr;

    fn check(rc: std::os::raw::c_int) -> Result<(), GpuError> {
        if rc == ffi::SCIR_B200_OK {
            return Ok(());
        }
        if rc == ffi::SCIR_B200_ERR_SHAPE {
            return Err(GpuError::ShapeMismatch);
        }
        let msg = unsafe { CStr::from_ptr(ffi::scir_b200_last_error()) }.to_string_lossy().into_owned();
        Err(GpuError::BackendUnavailable(format!("scir_b200 error {rc}: {msg}")))
    }

    fn ld(n: usize) -> i64 {
        if n > 0 { n as i64 } else { 1 }
    }

    /// Padding of [`filtfilt_fir`]: SciPy's `padtype` plus the reference's own unpadded structure.
    #[derive(Clone, Copy, Debug, PartialEq, Eq)]
    pub enum PadType {
        /// zero-state forward, reverse, zero-state forward, reverse (sig/lib.rs:278-291)
        ZeroState,
        /// SciPy default
        Odd,
        /// even extension
        Even,
        /// constant extension
        Constant,
        /// SciPy `padtype=None`
        NoPad,
    }

    impl PadType {
        fn code(self) -> std::os::raw::c_int {
            match self {
                PadType::ZeroState => ffi::SCIR_B200_PAD_ZERO_STATE,
                PadType::Odd => ffi::SCIR_B200_PAD_ODD,
                PadType::Even => ffi::SCIR_B200_PAD_EVEN,
                PadType::Constant => ffi::SCIR_B200_PAD_CONSTANT,
                PadType::NoPad => ffi::SCIR_B200_PAD_SCIPY_NONE,
            }
        }
    }

    /// `lfilter(b, [1], x)` along the rows: `y[r,i] = sum_d b[d] * x[r,i-d]` (SciPy tap order -- the reverse of
    /// `fir1d_batched_f32`'s).
    pub fn lfilter_fir(b: &Array1<f32>, x: &Array2<f32>) -> Result<Array2<f32>, GpuError> {
        let (rows, n) = x.dim();
        let xs = x.as_standard_layout();
        let bs = b.as_standard_layout();
        let mut out = vec![0.0f32; rows * n];
        with_default_context(|ctx| {
            check(unsafe {
                ffi::scir_b200_fir1d_batched_f32_host(
                    ctx.as_ptr(), xs.as_ptr(), ld(n), bs.as_ptr(), bs.len() as i64, ffi::SCIR_B200_TAPS_LFILTER,
                    out.as_mut_ptr(), ld(n), rows as i64, n as i64,
                )
            })
        })?;
        Array2::from_shape_vec((rows, n), out).map_err(|_| GpuError::ShapeMismatch)
    }

    /// `resample_poly(x, up, down, window=h)` along the rows, `padtype='constant'` (SciPy
    /// `_signaltools.py:3865-3957`); output is `(rows, ceil(n*up/down))`.
    pub fn resample_poly(x: &Array2<f32>, up: usize, down: usize, window: &Array1<f32>) -> Result<Array2<f32>, GpuError> {
        let (rows, n) = x.dim();
        let mut plan = ffi::ScirB200ResamplePlan::default();
        check(unsafe { ffi::scir_b200_resample_poly_plan(n as i64, window.len() as i64, up as i64, down as i64, &mut plan) })?;
        let n_out = if plan.up == 1 && plan.down == 1 { n } else { plan.n_out as usize };
        let xs = x.as_standard_layout();
        let ws = window.as_standard_layout();
        let mut out = vec![0.0f32; rows * n_out];
        with_default_context(|ctx| {
            check(unsafe {
                ffi::scir_b200_resample_poly_f32_host(
                    ctx.as_ptr(), ws.as_ptr(), ws.len() as i64, up as i64, down as i64, xs.as_ptr(), ld(n), rows as i64,
                    n as i64, out.as_mut_ptr(), ld(n_out),
                )
            })
        })?;
        Array2::from_shape_vec((rows, n_out), out).map_err(|_| GpuError::ShapeMismatch)
    }

    /// Zero-phase forward-backward filtering with an FIR numerator `b` (SciPy `filtfilt(b, [1], x, padtype, padlen)`,
    /// `_signaltools.py:4745-4826`; `PadType::ZeroState` is the reference's structure).  `padlen = None`: 3 * len(b).
    pub fn filtfilt_fir(b: &Array1<f32>, x: &Array2<f32>, pad: PadType, padlen: Option<usize>) -> Result<Array2<f32>, GpuError> {
        let (rows, n) = x.dim();
        let xs = x.as_standard_layout();
        let bs = b.as_standard_layout();
        let mut out = vec![0.0f32; rows * n];
        let pl = padlen.map(|p| p as i64).unwrap_or(-1);
        with_default_context(|ctx| {
            check(unsafe {
                ffi::scir_b200_filtfilt_fir_f32_host(
                    ctx.as_ptr(), bs.as_ptr(), bs.len() as i64, pad.code(), pl, xs.as_ptr(), ld(n), out.as_mut_ptr(), ld(n),
                    rows as i64, n as i64,
                )
            })
        })?;
        Array2::from_shape_vec((rows, n), out).map_err(|_| GpuError::ShapeMismatch)
    }

    /// Device scratch for the f64 routes (they take device pointers): RAII around `scir_b200_malloc` / `free`.
    struct DevBuf(*mut std::os::raw::c_void);

    impl DevBuf {
        fn new(ctx: &scir_gpu::Context, bytes: usize) -> Result<Self, GpuError> {
            let mut p = std::ptr::null_mut();
            check(unsafe { ffi::scir_b200_malloc(ctx.as_ptr(), bytes.max(8), &mut p) })?;
            Ok(DevBuf(p))
        }
    }

    fn run_f64(
        x: &[f64],
        n_out_total: usize,
        f: impl FnOnce(&scir_gpu::Context, *const f64, *mut f64) -> std::os::raw::c_int,
    ) -> Result<Vec<f64>, GpuError> {
        let mut out = vec![0.0f64; n_out_total];
        with_default_context(|ctx| {
            let dx = DevBuf::new(ctx, x.len() * 8)?;
            let dy = DevBuf::new(ctx, n_out_total * 8)?;
            let res = (|| {
                check(unsafe { ffi::scir_b200_memcpy_h2d(ctx.as_ptr(), dx.0, x.as_ptr() as *const _, x.len() * 8) })?;
                check(f(ctx, dx.0 as *const f64, dy.0 as *mut f64))?;
                check(unsafe { ffi::scir_b200_memcpy_d2h(ctx.as_ptr(), out.as_mut_ptr() as *mut _, dy.0, n_out_total * 8) })
            })();
            unsafe {
                ffi::scir_b200_free(ctx.as_ptr(), dx.0);
                ffi::scir_b200_free(ctx.as_ptr(), dy.0);
            }
            res
        })?;
        Ok(out)
    }

    fn run_f32(
        x: &[f32],
        n_out_total: usize,
        f: impl FnOnce(&scir_gpu::Context, *const f32, *mut f32) -> std::os::raw::c_int,
    ) -> Result<Vec<f32>, GpuError> {
        let mut out = vec![0.0f32; n_out_total];
        with_default_context(|ctx| {
            let dx = DevBuf::new(ctx, x.len() * 4)?;
            let dy = DevBuf::new(ctx, n_out_total * 4)?;
            let res = (|| {
                check(unsafe { ffi::scir_b200_memcpy_h2d(ctx.as_ptr(), dx.0, x.as_ptr() as *const _, x.len() * 4) })?;
                check(f(ctx, dx.0 as *const f32, dy.0 as *mut f32))?;
                check(unsafe { ffi::scir_b200_memcpy_d2h(ctx.as_ptr(), out.as_mut_ptr() as *mut _, dy.0, n_out_total * 4) })
            })();
            unsafe {
                ffi::scir_b200_free(ctx.as_ptr(), dx.0);
                ffi::scir_b200_free(ctx.as_ptr(), dy.0);
            }
            res
        })?;
        Ok(out)
    }

    /// Signal-extension modes of `upfirdn` / `resample_poly` (SciPy's MODE enum, `_upfirdn_apply.pyx:77-86`).
    #[derive(Clone, Copy, Debug, PartialEq, Eq)]
    pub enum ExtMode {
        /// pad with `cval`
        Constant,
        /// d c b a | a b c d | d c b a
        Symmetric,
        /// a a a a | a b c d | d d d d
        Edge,
        /// linear continuation of the first / last slope
        Smooth,
        /// a b c d | a b c d | a b c d
        Wrap,
        /// d c b | a b c d | c b a
        Reflect,
        /// -d -c -b -a | a b c d | -d -c -b -a
        Antisymmetric,
        /// odd reflection about the end samples
        Antireflect,
        /// the line through the first and last sample
        Line,
    }

    impl ExtMode {
        fn code(self) -> std::os::raw::c_int {
            match self {
                ExtMode::Constant => ffi::SCIR_B200_EXT_CONSTANT,
                ExtMode::Symmetric => ffi::SCIR_B200_EXT_SYMMETRIC,
                ExtMode::Edge => ffi::SCIR_B200_EXT_EDGE,
                ExtMode::Smooth => ffi::SCIR_B200_EXT_SMOOTH,
                ExtMode::Wrap => ffi::SCIR_B200_EXT_PERIODIC,
                ExtMode::Reflect => ffi::SCIR_B200_EXT_REFLECT,
                ExtMode::Antisymmetric => ffi::SCIR_B200_EXT_ANTISYMMETRIC,
                ExtMode::Antireflect => ffi::SCIR_B200_EXT_ANTIREFLECT,
                ExtMode::Line => ffi::SCIR_B200_EXT_LINE,
            }
        }
    }

    /// `upfirdn(h, x, up, down, mode, cval)` along the rows (SciPy `_upfirdn.py:107-216`); output is
    /// `(rows, _output_len(len(h), n, up, down))`.
    pub fn upfirdn(h: &Array1<f32>, x: &Array2<f32>, up: usize, down: usize, mode: ExtMode, cval: f32) -> Result<Array2<f32>, GpuError> {
        let (rows, n) = x.dim();
        let n_out = unsafe { ffi::scir_b200_upfirdn_out_len(h.len() as i64, n as i64, up as i64, down as i64) } as usize;
        let xs = x.as_standard_layout();
        let hs = h.as_standard_layout();
        let out = run_f32(xs.as_slice().ok_or(GpuError::ShapeMismatch)?, rows * n_out, |ctx, dx, dy| unsafe {
            ffi::scir_b200_upfirdn_mode_f32(
                ctx.as_ptr(), hs.as_ptr(), hs.len() as i64, up as i64, down as i64, mode.code(), cval, dx, ld(n), rows as i64,
                n as i64, dy, ld(n_out), 0, n_out as i64,
            )
        })?;
        Array2::from_shape_vec((rows, n_out), out).map_err(|_| GpuError::ShapeMismatch)
    }

    /// `resample_poly(x, up, down, window=h, padtype=<extension mode>, cval)` (SciPy `_signaltools.py:3921-3957`).
    pub fn resample_poly_pad(
        x: &Array2<f32>, up: usize, down: usize, window: &Array1<f32>, padtype: ExtMode, cval: f32,
    ) -> Result<Array2<f32>, GpuError> {
        let (rows, n) = x.dim();
        let mut plan = ffi::ScirB200ResamplePlan::default();
        check(unsafe { ffi::scir_b200_resample_poly_plan(n as i64, window.len() as i64, up as i64, down as i64, &mut plan) })?;
        let n_out = if plan.up == 1 && plan.down == 1 { n } else { plan.n_out as usize };
        let xs = x.as_standard_layout();
        let ws = window.as_standard_layout();
        let out = run_f32(xs.as_slice().ok_or(GpuError::ShapeMismatch)?, rows * n_out, |ctx, dx, dy| unsafe {
            ffi::scir_b200_resample_poly_pad_f32(
                ctx.as_ptr(), ws.as_ptr(), ws.len() as i64, up as i64, down as i64, padtype.code(), cval, dx, ld(n), rows as i64,
                n as i64, dy, ld(n_out),
            )
        })?;
        Array2::from_shape_vec((rows, n_out), out).map_err(|_| GpuError::ShapeMismatch)
    }

    /// f64 twin of the reference's `resample_poly(input: &Array1<f64>, up, down)` (sig/lib.rs:313-362) for ANY rate
    /// and filter: SciPy's `resample_poly(x, up, down, window=h)` on one f64 row, computed in f64 on the device.
    pub fn resample_poly_f64(input: &Array1<f64>, up: usize, down: usize, window: &Array1<f64>) -> Result<Array1<f64>, GpuError> {
        let n = input.len();
        let mut plan = ffi::ScirB200ResamplePlan::default();
        check(unsafe { ffi::scir_b200_resample_poly_plan(n as i64, window.len() as i64, up as i64, down as i64, &mut plan) })?;
        let n_out = if plan.up == 1 && plan.down == 1 { n } else { plan.n_out as usize };
        let xs = input.as_standard_layout();
        let ws = window.as_standard_layout();
        let out = run_f64(xs.as_slice().ok_or(GpuError::ShapeMismatch)?, n_out, |ctx, dx, dy| unsafe {
            ffi::scir_b200_resample_poly_pad_f64(
                ctx.as_ptr(), ws.as_ptr(), ws.len() as i64, up as i64, down as i64, ffi::SCIR_B200_EXT_CONSTANT, 0.0, dx,
                ld(n), 1, n as i64, dy, ld(n_out),
            )
        })?;
        Ok(Array1::from(out))
    }

    /// f64 zero-phase FIR filtering of one row (the FIR counterpart of the reference's SOS `filtfilt`, sig/lib.rs:278-291).
    pub fn filtfilt_fir_f64(b: &Array1<f64>, input: &Array1<f64>, pad: PadType, padlen: Option<usize>) -> Result<Array1<f64>, GpuError> {
        let n = input.len();
        let xs = input.as_standard_layout();
        let bs = b.as_standard_layout();
        let pl = padlen.map(|p| p as i64).unwrap_or(-1);
        let out = run_f64(xs.as_slice().ok_or(GpuError::ShapeMismatch)?, n, |ctx, dx, dy| unsafe {
            ffi::scir_b200_filtfilt_fir_f64(ctx.as_ptr(), bs.as_ptr(), bs.len() as i64, pad.code(), pl, dx, ld(n), dy, ld(n), 1, n as i64)
        })?;
        Ok(Array1::from(out))
    }
}

#[cfg(feature = "cuda")]
pub use routes::{
    filtfilt_fir, filtfilt_fir_f64, lfilter_fir, resample_poly, resample_poly_f64, resample_poly_pad, upfirdn, ExtMode, PadType,
};
