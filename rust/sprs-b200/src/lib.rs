//! Safe wrapper: a device-backed matrix with sprs's product call shapes.
//!
//! Orphan rules forbid re-implementing `Mul` for the foreign `CsMatBase`, so the
//! drop-in is a newtype owning the host `CsMatI` plus its device mirror:
//!
//! ```ignore
//! let a = DeviceCsMat::new(a_host)?;        // uploads once (proper_indptr + as_ptr)
//! let y: Array1<f64> = &a * &x;             // csmat.rs:2119-2160 -> sprs_b200_mul_mat_vec
//! let c: Array2<f64> = &a * &b;             // csmat.rs:1989-2048 (k >= 8 -> rowmaj, C order)
//! let p: CsMatI<f64, I, Iptr> = &a * &b_sp; // csmat.rs:1866-1949 -> smmp::mul_csr_csr
//! ```
//! Contract violations panic with the reference's messages (Guidelines.rst:9-27);
//! device failures are `LinalgError::ThirdPartyError(code, msg)` (errors.rs:70).
use ndarray::{Array1, Array2, ArrayBase, ArrayView1, Data, Ix1, Ix2, ShapeBuilder};
use sprs::errors::LinalgError;
use sprs::{CsMatI, CsMatViewI, CsVecI, SpIndex};
use sprs_b200_sys as ffi;
use std::ffi::CStr;
use std::ops::Mul;
use std::os::raw::c_void;

thread_local! { static CTX: Ctx = Ctx::new(0).expect("no B200 device"); }

struct Ctx(*mut ffi::sprs_b200_ctx);
impl Ctx {
    fn new(device: i32) -> Result<Self, LinalgError> {
        let mut h = std::ptr::null_mut();
        let st = unsafe { ffi::sprs_b200_ctx_create(device, &mut h) };
        if st != ffi::SPRS_B200_OK { return Err(third_party(std::ptr::null(), st)); }
        Ok(Ctx(h))
    }
}
impl Drop for Ctx { fn drop(&mut self) { unsafe { ffi::sprs_b200_ctx_destroy(self.0); } } }

fn third_party(ctx: *const ffi::sprs_b200_ctx, code: i32) -> LinalgError {
    // errors.rs:70 wants a &'static str; leak the (rare) message like a panic payload
    let msg = unsafe { CStr::from_ptr(ffi::sprs_b200_last_error(ctx)) }.to_string_lossy().into_owned();
    LinalgError::ThirdPartyError(code as isize, Box::leak(msg.into_boxed_str()))
}
fn check(ctx: *const ffi::sprs_b200_ctx, st: i32) -> Result<(), LinalgError> {
    match st {
        ffi::SPRS_B200_OK => Ok(()),
        ffi::SPRS_B200_ERR_DIMENSION => panic!("Dimension mismatch"),
        ffi::SPRS_B200_ERR_STORAGE => panic!("Storage mismatch"),
        _ => Err(third_party(ctx, st)),
    }
}

/// Host `CsMatI` + device mirror; the mirror is released in `Drop` (UMFPACK pattern,
/// sprs_suitesparse_umfpack/src/lib.rs:33-46).
pub struct DeviceCsMat<I: SpIndex, Iptr: SpIndex = I> {
    host: CsMatI<f64, I, Iptr>,
    dev: *mut ffi::sprs_b200_csmat,
}

impl<I: SpIndex, Iptr: SpIndex> DeviceCsMat<I, Iptr> {
    pub fn new(host: CsMatI<f64, I, Iptr>) -> Result<Self, LinalgError> {
        assert!(matches!(std::mem::size_of::<I>(), 4 | 8) && matches!(std::mem::size_of::<Iptr>(), 4 | 8),
                "device mirrors take 4- or 8-byte index types; use to_other_types() first");
        let dev = upload(host.view())?;
        Ok(Self { host, dev })
    }
    pub fn host(&self) -> &CsMatI<f64, I, Iptr> { &self.host }
}
impl<I: SpIndex, Iptr: SpIndex> Drop for DeviceCsMat<I, Iptr> {
    fn drop(&mut self) { unsafe { ffi::sprs_b200_csmat_free(self.dev); } }
}

fn upload<I: SpIndex, Iptr: SpIndex>(m: CsMatViewI<f64, I, Iptr>) -> Result<*mut ffi::sprs_b200_csmat, LinalgError> {
    // like the reference's own FFI callers (sprs-benches/src/main.rs:55-58): proper
    // (zero-based) indptr and raw as_ptr(); the library rebases anyway.
    let indptr = m.proper_indptr();
    let mut out = std::ptr::null_mut();
    CTX.with(|c| check(c.0, unsafe {
        ffi::sprs_b200_csmat_upload(
            c.0, if m.is_csr() { ffi::SPRS_B200_CSR } else { ffi::SPRS_B200_CSC },
            m.rows() as u64, m.cols() as u64,
            indptr.as_ptr() as *const c_void, std::mem::size_of::<Iptr>() as i32,
            m.indices().as_ptr() as *const c_void, std::mem::size_of::<I>() as i32,
            m.data().as_ptr(), &mut out)
    }))?;
    Ok(out)
}

// `&A * &x`  (csmat.rs:2119-2160)
impl<'a, 'b, I: SpIndex, Iptr: SpIndex, DS: Data<Elem = f64>> Mul<&'b ArrayBase<DS, Ix1>> for &'a DeviceCsMat<I, Iptr> {
    type Output = Array1<f64>;
    fn mul(self, rhs: &'b ArrayBase<DS, Ix1>) -> Array1<f64> {
        assert_eq!(self.host.cols(), rhs.len(), "Dimension mismatch");
        let x = rhs.as_standard_layout();
        let mut y = Array1::<f64>::zeros(self.host.rows());
        CTX.with(|c| check(c.0, unsafe {
            ffi::sprs_b200_mul_mat_vec(c.0, self.dev, x.as_ptr(), x.len() as u64,
                                       y.as_mut_ptr(), y.len() as u64)
        })).expect("sprs_b200 device error");
        y
    }
}

// `&A * &v`, v sparse (vec.rs:1104-1131 -> prod::csr_mul_csvec, prod.rs:162-184): the per-row
// sorted-merge dot runs on the device (bit-identical); zeros are dropped here (prod.rs:178-180).
// A CSC lhs goes through the sparse-sparse product with `rhs.col_view()` like the reference.
impl<'a, 'b, I: SpIndex, Iptr: SpIndex> Mul<&'b CsVecI<f64, I>> for &'a DeviceCsMat<I, Iptr> {
    type Output = CsVecI<f64, I>;
    fn mul(self, rhs: &'b CsVecI<f64, I>) -> CsVecI<f64, I> {
        if rhs.dim() == 0 { return CsVecI::empty(0); }
        assert_eq!(self.host.cols(), rhs.dim(), "Dimension mismatch");
        assert!(self.host.is_csr(), "CSC x CsVec: use &a * &DeviceCsMat::new(rhs.col_view().to_owned())");
        let mut y = vec![0f64; self.host.rows()];
        CTX.with(|c| check(c.0, unsafe {
            ffi::sprs_b200_csr_mul_csvec(c.0, self.dev, rhs.dim() as u64, rhs.nnz() as u64,
                rhs.indices().as_ptr() as *const c_void, std::mem::size_of::<I>() as i32,
                rhs.data().as_ptr(), y.as_mut_ptr(), y.len() as u64)
        })).expect("sprs_b200 device error");
        let mut res = CsVecI::empty(self.host.rows());
        for (row, val) in y.into_iter().enumerate() {
            if val != 0.0 { res.append(row, val); }
        }
        res
    }
}

// `&A * &B`, dense B (csmat.rs:1989-2048): k >= 8 -> rowmaj kernel, C-order result
impl<'a, 'b, I: SpIndex, Iptr: SpIndex, DS: Data<Elem = f64>> Mul<&'b ArrayBase<DS, Ix2>> for &'a DeviceCsMat<I, Iptr> {
    type Output = Array2<f64>;
    fn mul(self, rhs: &'b ArrayBase<DS, Ix2>) -> Array2<f64> {
        let (rows, cols) = (self.host.rows(), rhs.shape()[1]);
        let (rs, cs) = (rhs.strides()[0] as i64, rhs.strides()[1] as i64);
        let wide = cols >= 8; // csmat.rs:2009
        let mut res = if wide { Array2::zeros((rows, cols)) } else { Array2::zeros((rows, cols).f()) };
        let (ors, ocs) = (res.strides()[0] as i64, res.strides()[1] as i64);
        let f = match (self.host.is_csr(), wide) {
            (true, true) => ffi::sprs_b200_csr_mulacc_dense_rowmaj,
            (true, false) => ffi::sprs_b200_csr_mulacc_dense_colmaj,
            (false, true) => ffi::sprs_b200_csc_mulacc_dense_rowmaj,
            (false, false) => ffi::sprs_b200_csc_mulacc_dense_colmaj,
        };
        CTX.with(|c| check(c.0, unsafe {
            f(c.0, self.dev, rhs.as_ptr(), rhs.shape()[0] as u64, cols as u64, rs, cs,
              res.as_mut_ptr(), rows as u64, cols as u64, ors, ocs)
        })).expect("sprs_b200 device error");
        res
    }
}

/// smmp::mul_csr_csr (smmp.rs:196-237): Rust allocates the output Vecs between the
/// symbolic and numeric calls, exactly where the reference does.
pub fn mul_csr_csr<I: SpIndex, Iptr: SpIndex>(lhs: &DeviceCsMat<I, Iptr>, rhs: &DeviceCsMat<I, Iptr>) -> CsMatI<f64, I, Iptr> {
    assert_eq!(lhs.host.cols(), rhs.host.rows());
    assert!(lhs.host.is_csr() && rhs.host.is_csr(), "Storage mismatch");
    CTX.with(|c| {
        let (mut plan, mut nnz_c) = (std::ptr::null_mut(), 0u64);
        check(c.0, unsafe { ffi::sprs_b200_spgemm_symbolic(c.0, lhs.dev, rhs.dev, &mut plan, &mut nnz_c) })
            .expect("sprs_b200 device error");
        let mut indptr = vec![Iptr::zero(); lhs.host.rows() + 1];
        let mut indices = vec![I::zero(); nnz_c as usize];
        let mut data = vec![0f64; nnz_c as usize];
        let st = unsafe {
            ffi::sprs_b200_spgemm_numeric(c.0, plan, indptr.as_mut_ptr() as *mut c_void,
                std::mem::size_of::<Iptr>() as i32, indices.as_mut_ptr() as *mut c_void,
                std::mem::size_of::<I>() as i32, data.as_mut_ptr())
        };
        unsafe { ffi::sprs_b200_spgemm_free(plan); }
        check(c.0, st).expect("sprs_b200 device error");
        // invariants hold by construction (sorted unique in-range columns): smmp.rs:406-415
        CsMatI::new_trusted(sprs::CompressedStorage::CSR, (lhs.host.rows(), rhs.host.cols()), indptr, indices, data)
    })
}

// `&A * &B`, both sparse (csmat.rs:1866-1949) for the (CSR, CSR) case; the mixed-storage
// arms convert with to_other_storage() exactly as csmat_mul_csmat does.
impl<'a, 'b, I: SpIndex, Iptr: SpIndex> Mul<&'b DeviceCsMat<I, Iptr>> for &'a DeviceCsMat<I, Iptr> {
    type Output = CsMatI<f64, I, Iptr>;
    fn mul(self, rhs: &'b DeviceCsMat<I, Iptr>) -> Self::Output { mul_csr_csr(self, rhs) }
}

/// sprs::linalg::bicgstab::BiCGSTAB<f64> (linalg/bicgstab.rs:95-300) with x, r, rhat, p
/// resident on the device between iterations.  Same constructor, `solve`, `step`,
/// restarts and accessors; vectors cross the API as dense `Array1<f64>`.
pub mod bicgstab {
    use super::*;
    pub struct BiCGSTAB<'a, I: SpIndex, Iptr: SpIndex> { a: &'a DeviceCsMat<I, Iptr>, h: *mut ffi::sprs_b200_bicgstab }
    impl<'a, I: SpIndex, Iptr: SpIndex> Drop for BiCGSTAB<'a, I, Iptr> {
        fn drop(&mut self) { unsafe { ffi::sprs_b200_bicgstab_free(self.h); } }
    }
    impl<'a, I: SpIndex, Iptr: SpIndex> BiCGSTAB<'a, I, Iptr> {
        /// bicgstab.rs:120-146
        pub fn new(a: &'a DeviceCsMat<I, Iptr>, x0: ArrayView1<f64>, b: ArrayView1<f64>) -> Self {
            assert_eq!(a.host.cols(), x0.len(), "Dimension mismatch");
            assert_eq!(a.host.rows(), b.len(), "Dimension mismatch");
            let (x0, b) = (x0.to_owned(), b.to_owned()); // contiguous
            let mut h = std::ptr::null_mut();
            CTX.with(|c| check(c.0, unsafe {
                ffi::sprs_b200_bicgstab_new(c.0, a.dev, x0.as_ptr(), b.as_ptr(), b.len() as u64, &mut h)
            })).expect("sprs_b200 device error");
            Self { a, h }
        }
        /// bicgstab.rs:151-175
        pub fn solve(a: &'a DeviceCsMat<I, Iptr>, x0: ArrayView1<f64>, b: ArrayView1<f64>, tol: f64,
                     max_iter: usize) -> Result<Box<Self>, Box<Self>> {
            let solver = Self::new(a, x0, b);
            let mut converged = 0;
            CTX.with(|c| check(c.0, unsafe {
                ffi::sprs_b200_bicgstab_solve(solver.h, tol, max_iter as u64, &mut converged)
            })).expect("sprs_b200 device error");
            if converged != 0 { Ok(Box::new(solver)) } else { Err(Box::new(solver)) }
        }
        pub fn step(&mut self) -> f64 {
            let mut err = 0.0;
            CTX.with(|c| check(c.0, unsafe { ffi::sprs_b200_bicgstab_step(self.h, &mut err) }))
                .expect("sprs_b200 device error");
            err
        }
        pub fn soft_restart(&mut self) { unsafe { ffi::sprs_b200_bicgstab_soft_restart(self.h); } }
        pub fn hard_restart(&mut self) { unsafe { ffi::sprs_b200_bicgstab_hard_restart(self.h); } }
        pub fn with_restart_threshold(self, thresh: f64) -> Self {
            unsafe { ffi::sprs_b200_bicgstab_set_restart_threshold(self.h, thresh); }
            self
        }
        fn stats(&self) -> ([u64; 3], [f64; 3]) {
            let (mut c, mut s) = ([0u64; 3], [0f64; 3]);
            unsafe { ffi::sprs_b200_bicgstab_stats(self.h, c.as_mut_ptr(), s.as_mut_ptr()); }
            (c, s)
        }
        pub fn iteration_count(&self) -> usize { self.stats().0[0] as usize }
        pub fn soft_restart_count(&self) -> usize { self.stats().0[1] as usize }
        pub fn hard_restart_count(&self) -> usize { self.stats().0[2] as usize }
        pub fn err(&self) -> f64 { self.stats().1[0] }
        pub fn rho(&self) -> f64 { self.stats().1[1] }
        pub fn soft_restart_threshold(&self) -> f64 { self.stats().1[2] }
        pub fn a(&self) -> &DeviceCsMat<I, Iptr> { self.a }
        fn vec(&self, which: i32) -> Array1<f64> {
            let mut out = Array1::zeros(self.a.host.rows());
            unsafe { ffi::sprs_b200_bicgstab_get(self.h, which, out.as_mut_ptr(), out.len() as u64); }
            out
        }
        pub fn x(&self) -> Array1<f64> { self.vec(ffi::SPRS_B200_BICGSTAB_X) }
        pub fn b(&self) -> Array1<f64> { self.vec(ffi::SPRS_B200_BICGSTAB_B) }
        pub fn r(&self) -> Array1<f64> { self.vec(ffi::SPRS_B200_BICGSTAB_R) }
        pub fn rhat(&self) -> Array1<f64> { self.vec(ffi::SPRS_B200_BICGSTAB_RHAT) }
        pub fn p(&self) -> Array1<f64> { self.vec(ffi::SPRS_B200_BICGSTAB_P) }
    }
}
