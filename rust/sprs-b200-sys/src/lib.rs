//! Raw declarations of include/sprs_b200.h.  Conventions follow the in-tree FFI
//! precedent `prod_nnz` (sprs-benches/src/main.rs:27-42): plain scalars, raw pointers,
//! zero-based indptr, caller-owned host buffers; opaque handles freed explicitly
//! (suitesparse_umfpack_sys/src/umfpack_free_numeric.rs:3-6).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)] pub struct sprs_b200_ctx { _private: [u8; 0] }
#[repr(C)] pub struct sprs_b200_csmat { _private: [u8; 0] }
#[repr(C)] pub struct sprs_b200_spgemm { _private: [u8; 0] }
#[repr(C)] pub struct sprs_b200_bicgstab { _private: [u8; 0] }

pub const SPRS_B200_CSR: c_int = 0;
pub const SPRS_B200_CSC: c_int = 1;
pub const SPRS_B200_OK: c_int = 0;
pub const SPRS_B200_ERR_DIMENSION: c_int = 1;
pub const SPRS_B200_ERR_STORAGE: c_int = 2;
pub const SPRS_B200_ERR_CUDA: c_int = 3;
pub const SPRS_B200_ERR_NCCL: c_int = 4;
pub const SPRS_B200_ERR_INDEX_RANGE: c_int = 5;
pub const SPRS_B200_BICGSTAB_X: c_int = 0;
pub const SPRS_B200_BICGSTAB_R: c_int = 1;
pub const SPRS_B200_BICGSTAB_RHAT: c_int = 2;
pub const SPRS_B200_BICGSTAB_P: c_int = 3;
pub const SPRS_B200_BICGSTAB_B: c_int = 4;

/// y = A x callback of the operator-form solver: device pointers to n doubles, a cudaStream_t.
pub type sprs_b200_matvec_fn = Option<unsafe extern "C" fn(
    user: *mut c_void, d_x: *const c_double, d_y: *mut c_double, stream: *mut c_void) -> c_int>;

extern "C" {
    pub fn sprs_b200_version() -> c_int;
    pub fn sprs_b200_ctx_create(device: c_int, out: *mut *mut sprs_b200_ctx) -> c_int;
    pub fn sprs_b200_ctx_destroy(ctx: *mut sprs_b200_ctx) -> c_int;
    pub fn sprs_b200_last_error(ctx: *const sprs_b200_ctx) -> *const c_char;
    pub fn sprs_b200_csmat_upload(
        ctx: *mut sprs_b200_ctx, storage: c_int, rows: u64, cols: u64,
        indptr: *const c_void, indptr_bytes: c_int,
        indices: *const c_void, index_bytes: c_int,
        data: *const c_double, out: *mut *mut sprs_b200_csmat) -> c_int;
    pub fn sprs_b200_csmat_free(m: *mut sprs_b200_csmat) -> c_int;
    pub fn sprs_b200_csmat_nnz(m: *const sprs_b200_csmat) -> u64;
    pub fn sprs_b200_csmat_download(
        ctx: *mut sprs_b200_ctx, m: *const sprs_b200_csmat, indptr: *mut c_void,
        indptr_bytes: c_int, indices: *mut c_void, index_bytes: c_int, data: *mut c_double) -> c_int;
    pub fn sprs_b200_csmat_to_other_storage(
        ctx: *mut sprs_b200_ctx, m: *const sprs_b200_csmat, out: *mut *mut sprs_b200_csmat) -> c_int;
    pub fn sprs_b200_mul_acc_mat_vec_csr(
        ctx: *mut sprs_b200_ctx, mat: *const sprs_b200_csmat, in_vec: *const c_double, in_len: u64,
        res_vec: *mut c_double, res_len: u64) -> c_int;
    pub fn sprs_b200_mul_acc_mat_vec_csc(
        ctx: *mut sprs_b200_ctx, mat: *const sprs_b200_csmat, in_vec: *const c_double, in_len: u64,
        res_vec: *mut c_double, res_len: u64) -> c_int;
    pub fn sprs_b200_mul_mat_vec(
        ctx: *mut sprs_b200_ctx, mat: *const sprs_b200_csmat, x: *const c_double, x_len: u64,
        y: *mut c_double, y_len: u64) -> c_int;
    pub fn sprs_b200_csr_mul_csvec(
        ctx: *mut sprs_b200_ctx, mat: *const sprs_b200_csmat, dim: u64, v_nnz: u64,
        v_indices: *const c_void, index_bytes: c_int, v_data: *const c_double,
        res: *mut c_double, res_len: u64) -> c_int;
    pub fn sprs_b200_csr_mulacc_dense_rowmaj(
        ctx: *mut sprs_b200_ctx, lhs: *const sprs_b200_csmat, rhs: *const c_double,
        rhs_rows: u64, rhs_cols: u64, rhs_rs: i64, rhs_cs: i64, out: *mut c_double,
        out_rows: u64, out_cols: u64, out_rs: i64, out_cs: i64) -> c_int;
    pub fn sprs_b200_csr_mulacc_dense_colmaj(
        ctx: *mut sprs_b200_ctx, lhs: *const sprs_b200_csmat, rhs: *const c_double,
        rhs_rows: u64, rhs_cols: u64, rhs_rs: i64, rhs_cs: i64, out: *mut c_double,
        out_rows: u64, out_cols: u64, out_rs: i64, out_cs: i64) -> c_int;
    pub fn sprs_b200_csc_mulacc_dense_rowmaj(
        ctx: *mut sprs_b200_ctx, lhs: *const sprs_b200_csmat, rhs: *const c_double,
        rhs_rows: u64, rhs_cols: u64, rhs_rs: i64, rhs_cs: i64, out: *mut c_double,
        out_rows: u64, out_cols: u64, out_rs: i64, out_cs: i64) -> c_int;
    pub fn sprs_b200_csc_mulacc_dense_colmaj(
        ctx: *mut sprs_b200_ctx, lhs: *const sprs_b200_csmat, rhs: *const c_double,
        rhs_rows: u64, rhs_cols: u64, rhs_rs: i64, rhs_cs: i64, out: *mut c_double,
        out_rows: u64, out_cols: u64, out_rs: i64, out_cs: i64) -> c_int;
    pub fn sprs_b200_spgemm_symbolic(
        ctx: *mut sprs_b200_ctx, a: *const sprs_b200_csmat, b: *const sprs_b200_csmat,
        plan: *mut *mut sprs_b200_spgemm, nnz_c: *mut u64) -> c_int;
    pub fn sprs_b200_spgemm_numeric(
        ctx: *mut sprs_b200_ctx, plan: *mut sprs_b200_spgemm, c_indptr: *mut c_void,
        indptr_bytes: c_int, c_indices: *mut c_void, index_bytes: c_int, c_data: *mut c_double) -> c_int;
    pub fn sprs_b200_spgemm_free(plan: *mut sprs_b200_spgemm) -> c_int;
    // linalg::bicgstab::BiCGSTAB<f64> with device-resident vectors (bicgstab.rs:95-300)
    pub fn sprs_b200_bicgstab_new(
        ctx: *mut sprs_b200_ctx, mat: *const sprs_b200_csmat, x0: *const c_double,
        b: *const c_double, n: u64, out: *mut *mut sprs_b200_bicgstab) -> c_int;
    pub fn sprs_b200_bicgstab_new_op(
        ctx: *mut sprs_b200_ctx, n: u64, matvec: sprs_b200_matvec_fn, user: *mut c_void,
        x0: *const c_double, b: *const c_double, device_pointers: c_int,
        out: *mut *mut sprs_b200_bicgstab) -> c_int;
    pub fn sprs_b200_bicgstab_free(s: *mut sprs_b200_bicgstab) -> c_int;
    pub fn sprs_b200_bicgstab_step(s: *mut sprs_b200_bicgstab, err_out: *mut c_double) -> c_int;
    pub fn sprs_b200_bicgstab_soft_restart(s: *mut sprs_b200_bicgstab) -> c_int;
    pub fn sprs_b200_bicgstab_hard_restart(s: *mut sprs_b200_bicgstab) -> c_int;
    pub fn sprs_b200_bicgstab_solve(
        s: *mut sprs_b200_bicgstab, tol: c_double, max_iter: u64, converged: *mut c_int) -> c_int;
    pub fn sprs_b200_bicgstab_set_restart_threshold(s: *mut sprs_b200_bicgstab, thresh: c_double) -> c_int;
    pub fn sprs_b200_bicgstab_stats(
        s: *const sprs_b200_bicgstab, counts: *mut u64, scalars: *mut c_double) -> c_int;
    pub fn sprs_b200_bicgstab_get(
        s: *const sprs_b200_bicgstab, which: c_int, out: *mut c_double, len: u64) -> c_int;
}
