//! SOURCE ONLY (never compiled: no Rust toolchain in the build image).
//!
//! `impl Fft for GpuFft<T>`: the reference's operator interface (`fourier-algorithms/src/fft.rs:40-82`)
//! over the C ABI of libfourier.so (`include/fourier.h`), plus drop-in `create_fft_f32/f64`
//! (`fourier/src/lib.rs:31-60`).  Slices may live in host memory (staged over PCIe by the library).

use fourier_algorithms::{Fft, Transform};
use libc::{c_int, c_void, size_t};
use num_complex::Complex;

#[link(name = "fourier")]
extern "C" {
    fn fourier_create_float(size: size_t) -> *mut c_void;
    fn fourier_create_double(size: size_t) -> *mut c_void;
    fn fourier_destroy_float(plan: *mut c_void);
    fn fourier_destroy_double(plan: *mut c_void);
    fn fourier_transform_in_place_float(plan: *const c_void, data: *mut Complex<f32>, transform: c_int);
    fn fourier_transform_in_place_double(plan: *const c_void, data: *mut Complex<f64>, transform: c_int);
    fn fourier_transform_float(plan: *const c_void, i: *const Complex<f32>, o: *mut Complex<f32>, t: c_int);
    fn fourier_transform_double(plan: *const c_void, i: *const Complex<f64>, o: *mut Complex<f64>, t: c_int);
    // additive batched entry (include/fourier_b200.h)
    fn fourier_b200_transform_batch_float(p: *const c_void, i: *const c_void, o: *mut c_void, batch: size_t, t: c_int) -> c_int;
    fn fourier_b200_transform_batch_double(p: *const c_void, i: *const c_void, o: *mut c_void, batch: size_t, t: c_int) -> c_int;
}

/// Same codes as `convert_transform` in fourier-ffi/src/lib.rs:3-12.
fn code(t: Transform) -> c_int {
    match t {
        Transform::Fft => 0,
        Transform::Ifft => 1,
        Transform::UnscaledIfft => 2,
        Transform::SqrtScaledFft => 3,
        Transform::SqrtScaledIfft => 4,
    }
}

pub struct GpuFft32 { plan: *mut c_void, size: usize }
pub struct GpuFft64 { plan: *mut c_void, size: usize }
// The plan may move between threads but must not be shared (same contract as the reference's
// RefCell work buffer: autosort/mod.rs:54,151).
unsafe impl Send for GpuFft32 {}
unsafe impl Send for GpuFft64 {}

impl Fft for GpuFft32 {
    type Real = f32;
    fn size(&self) -> usize { self.size }
    fn transform_in_place(&self, input: &mut [Complex<f32>], transform: Transform) {
        assert_eq!(input.len(), self.size);
        unsafe { fourier_transform_in_place_float(self.plan, input.as_mut_ptr(), code(transform)) }
    }
    fn transform(&self, input: &[Complex<f32>], output: &mut [Complex<f32>], transform: Transform) {
        assert_eq!(input.len(), self.size);
        assert_eq!(output.len(), self.size);
        unsafe { fourier_transform_float(self.plan, input.as_ptr(), output.as_mut_ptr(), code(transform)) }
    }
}
impl Fft for GpuFft64 {
    type Real = f64;
    fn size(&self) -> usize { self.size }
    fn transform_in_place(&self, input: &mut [Complex<f64>], transform: Transform) {
        assert_eq!(input.len(), self.size);
        unsafe { fourier_transform_in_place_double(self.plan, input.as_mut_ptr(), code(transform)) }
    }
    fn transform(&self, input: &[Complex<f64>], output: &mut [Complex<f64>], transform: Transform) {
        assert_eq!(input.len(), self.size);
        assert_eq!(output.len(), self.size);
        unsafe { fourier_transform_double(self.plan, input.as_ptr(), output.as_mut_ptr(), code(transform)) }
    }
}
impl GpuFft32 {
    /// `batch` contiguous transforms (`data.len() == batch * size()`), in place.
    pub fn transform_batch_in_place(&self, data: &mut [Complex<f32>], transform: Transform) -> Result<(), i32> {
        assert_eq!(data.len() % self.size, 0);
        let p = data.as_mut_ptr() as *mut c_void;
        let rc = unsafe { fourier_b200_transform_batch_float(self.plan, p, p, data.len() / self.size, code(transform)) };
        if rc == 0 { Ok(()) } else { Err(rc) }
    }
}
impl GpuFft64 {
    pub fn transform_batch_in_place(&self, data: &mut [Complex<f64>], transform: Transform) -> Result<(), i32> {
        assert_eq!(data.len() % self.size, 0);
        let p = data.as_mut_ptr() as *mut c_void;
        let rc = unsafe { fourier_b200_transform_batch_double(self.plan, p, p, data.len() / self.size, code(transform)) };
        if rc == 0 { Ok(()) } else { Err(rc) }
    }
}
impl Drop for GpuFft32 { fn drop(&mut self) { unsafe { fourier_destroy_float(self.plan) } } }
impl Drop for GpuFft64 { fn drop(&mut self) { unsafe { fourier_destroy_double(self.plan) } } }

/// Drop-in for `fourier::create_fft_f32` (fourier/src/lib.rs:31-43).  Panics if no GPU plan can be built
/// (there is no CPU fallback in libfourier.so).
pub fn create_fft_f32(size: usize) -> Box<dyn Fft<Real = f32> + Send> {
    let plan = unsafe { fourier_create_float(size) };
    assert!(!plan.is_null(), "fourier_create_float returned NULL");
    Box::new(GpuFft32 { plan, size })
}
/// Drop-in for `fourier::create_fft_f64` (fourier/src/lib.rs:49-60).
pub fn create_fft_f64(size: usize) -> Box<dyn Fft<Real = f64> + Send> {
    let plan = unsafe { fourier_create_double(size) };
    assert!(!plan.is_null(), "fourier_create_double returned NULL");
    Box::new(GpuFft64 { plan, size })
}
