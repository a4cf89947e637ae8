// sprs_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement (C++17, host only) of the sprs product hot path, used as the
// parity oracle for the CUDA kernels in sprs_b200/csrc and as the "port" CPU
// baseline in bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product path
// (libsprs_b200.so) never links, loads or calls anything in oracle/.
//
// Why a restatement: the reference is Rust and this image has no cargo/rustc
// (SURVEY.md F1), so oracle/_ref cannot be built.  Parity is PINNED by the
// reference's own known-answer tests, transcribed as data into
// tests/golden/sprs_fixtures.json and checked by tests/test_oracle_golden.py:
//   sprs/src/test_data.rs:6-123, sprs/src/sparse/prod.rs:326-595,
//   sprs/src/sparse/smmp.rs:423-513, sprs/src/lib.rs:54-73.
//
// Arithmetic contract (sprs/src/mul_acc.rs:23-31): `*self += a * b` -- the
// product is rounded, then the sum is rounded; summation is strictly sequential
// in storage order.  Build with -ffp-contract=off so gcc never fuses to FMA.
//
// Every function cites the reference lines it follows.  Index types: I is the
// width of `indices`, P the width of `indptr` (sprs: I / Iptr, sparse.rs:94-109).
// `indptr` may be non-zero-based (row-sliced views, indptr.rs:122-124): like
// sprs, every access subtracts indptr[0].

#include <algorithm>
#include <cmath>
#include <functional>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------
// prod.rs:103-127  mul_acc_mat_vec_csr : res_vec[row] += A[row,:] . in_vec
// (identical arithmetic/order to csr_mulacc_dense_colmaj with one column, which
// is what `&A * &x` runs: csmat.rs:2142-2148, prod.rs:274-298).
template <class I, class P>
void mul_acc_mat_vec_csr(size_t rows, const P* indptr, const I* indices,
                         const double* data, const double* x, double* y) {
    const P base = indptr[0];
    for (size_t r = 0; r < rows; ++r) {
        double tv = y[r];
        for (size_t k = indptr[r] - base, e = indptr[r + 1] - base; k < e; ++k) {
            const double prod = data[k] * x[indices[k]];  // rounded product
            tv = tv + prod;                               // rounded sum (no FMA)
        }
        y[r] = tv;
    }
}

// prod.rs:74-99  mul_acc_mat_vec_csc : scatter-add, column by column.
template <class I, class P>
void mul_acc_mat_vec_csc(size_t cols, const P* indptr, const I* indices,
                         const double* data, const double* x, double* y) {
    const P base = indptr[0];
    for (size_t c = 0; c < cols; ++c) {
        const double xv = x[c];
        for (size_t k = indptr[c] - base, e = indptr[c + 1] - base; k < e; ++k) {
            const double prod = data[k] * xv;
            y[indices[k]] = y[indices[k]] + prod;
        }
    }
}

// prod.rs:274-298  csr_mulacc_dense_colmaj : for each rhs column, SpMV.
// rhs / out are ndarray views => arbitrary element strides (rs, cs).
template <class I, class P>
void csr_mulacc_dense_colmaj(size_t rows, size_t k_cols, const P* indptr,
                             const I* indices, const double* data,
                             const double* rhs, ptrdiff_t rhs_rs, ptrdiff_t rhs_cs,
                             double* out, ptrdiff_t out_rs, ptrdiff_t out_cs) {
    const P base = indptr[0];
    for (size_t c = 0; c < k_cols; ++c) {
        const double* rcol = rhs + (ptrdiff_t)c * rhs_cs;
        double* ocol = out + (ptrdiff_t)c * out_cs;
        for (size_t r = 0; r < rows; ++r) {
            double oval = ocol[(ptrdiff_t)r * out_rs];
            for (size_t k = indptr[r] - base, e = indptr[r + 1] - base; k < e; ++k) {
                const double prod = data[k] * rcol[(ptrdiff_t)indices[k] * rhs_rs];
                oval = oval + prod;
            }
            ocol[(ptrdiff_t)r * out_rs] = oval;
        }
    }
}

// prod.rs:189-214  csr_mulacc_dense_rowmaj : per nnz, k-wide axpy of a rhs row.
template <class I, class P>
void csr_mulacc_dense_rowmaj(size_t rows, size_t k_cols, const P* indptr,
                             const I* indices, const double* data,
                             const double* rhs, ptrdiff_t rhs_rs, ptrdiff_t rhs_cs,
                             double* out, ptrdiff_t out_rs, ptrdiff_t out_cs) {
    const P base = indptr[0];
    for (size_t r = 0; r < rows; ++r) {
        double* oline = out + (ptrdiff_t)r * out_rs;
        for (size_t k = indptr[r] - base, e = indptr[r + 1] - base; k < e; ++k) {
            const double lval = data[k];
            const double* rline = rhs + (ptrdiff_t)indices[k] * rhs_rs;
            for (size_t c = 0; c < k_cols; ++c) {
                const double prod = lval * rline[(ptrdiff_t)c * rhs_cs];
                oline[(ptrdiff_t)c * out_cs] = oline[(ptrdiff_t)c * out_cs] + prod;
            }
        }
    }
}

// prod.rs:219-241  csc_mulacc_dense_rowmaj
template <class I, class P>
void csc_mulacc_dense_rowmaj(size_t cols, size_t k_cols, const P* indptr,
                             const I* indices, const double* data,
                             const double* rhs, ptrdiff_t rhs_rs, ptrdiff_t rhs_cs,
                             double* out, ptrdiff_t out_rs, ptrdiff_t out_cs) {
    const P base = indptr[0];
    for (size_t lc = 0; lc < cols; ++lc) {
        const double* rline = rhs + (ptrdiff_t)lc * rhs_rs;
        for (size_t k = indptr[lc] - base, e = indptr[lc + 1] - base; k < e; ++k) {
            double* oline = out + (ptrdiff_t)indices[k] * out_rs;
            const double lval = data[k];
            for (size_t c = 0; c < k_cols; ++c) {
                const double prod = lval * rline[(ptrdiff_t)c * rhs_cs];
                oline[(ptrdiff_t)c * out_cs] = oline[(ptrdiff_t)c * out_cs] + prod;
            }
        }
    }
}

// prod.rs:246-269  csc_mulacc_dense_colmaj
template <class I, class P>
void csc_mulacc_dense_colmaj(size_t cols, size_t k_cols, const P* indptr,
                             const I* indices, const double* data,
                             const double* rhs, ptrdiff_t rhs_rs, ptrdiff_t rhs_cs,
                             double* out, ptrdiff_t out_rs, ptrdiff_t out_cs) {
    const P base = indptr[0];
    for (size_t c = 0; c < k_cols; ++c) {
        const double* rcol = rhs + (ptrdiff_t)c * rhs_cs;
        double* ocol = out + (ptrdiff_t)c * out_cs;
        for (size_t rrow = 0; rrow < cols; ++rrow) {
            const double rval = rcol[(ptrdiff_t)rrow * rhs_rs];
            for (size_t k = indptr[rrow] - base, e = indptr[rrow + 1] - base; k < e; ++k) {
                const double prod = data[k] * rval;
                double* o = ocol + (ptrdiff_t)indices[k] * out_rs;
                *o = *o + prod;
            }
        }
    }
}

// prod.rs:162-184 csr_mul_csvec + vec.rs:846-881 dot_acc (sorted two-pointer
// merge).  Exact zeros are dropped (prod.rs:179).  Returns nnz of the result.
template <class I, class P>
size_t csr_mul_csvec(size_t rows, const P* indptr, const I* indices,
                     const double* data, size_t v_nnz, const I* v_indices,
                     const double* v_data, I* out_indices, double* out_data) {
    const P base = indptr[0];
    size_t n_out = 0;
    for (size_t r = 0; r < rows; ++r) {
        size_t l = indptr[r] - base, le = indptr[r + 1] - base, q = 0;
        double sum = 0.0;
        while (l < le && q < v_nnz) {
            const I li = indices[l], ri = v_indices[q];
            if (li == ri) {
                const double prod = data[l] * v_data[q];
                sum = sum + prod;
            }
            if (li <= ri) ++l;
            if (li >= ri) ++q;
        }
        if (sum != 0.0) {
            out_indices[n_out] = (I)r;
            out_data[n_out] = sum;
            ++n_out;
        }
    }
    return n_out;
}

// prod.rs:13-72 csvec_dot_by_binary_search(+_impl): walk the vector with fewer entries,
// binary-search each index in the (shrinking) tail of the other, mul_acc on a match.
// The operand order of the product is preserved by the closure (prod.rs:25-32).
template <class I>
double csvec_dot_by_binary_search(size_t n1, const I* i1, const double* d1, size_t n2,
                                  const I* i2, const double* d2) {
    const bool swapped = n1 > n2;  // vec1.nnz() > vec2.nnz(): search in vec1 instead
    if (swapped) {
        std::swap(n1, n2);
        std::swap(i1, i2);
        std::swap(d1, d2);
    }
    double sum = 0.0;
    while (n1 != 0 && n2 != 0) {
        // slice::binary_search: Ok(i) when found, Err(i) = insertion point
        size_t lo = 0, hi = n2;
        bool found = false;
        while (lo < hi) {
            const size_t mid = lo + (hi - lo) / 2;
            if (i2[mid] == i1[0]) {
                found = true;
                lo = mid;
                break;
            }
            if (i2[mid] < i1[0]) lo = mid + 1;
            else hi = mid;
        }
        if (found) {
            const double prod = swapped ? d2[lo] * d1[0] : d1[0] * d2[lo];
            sum = sum + prod;
        }
        ++i1, ++d1, --n1;
        i2 += lo, d2 += lo, n2 -= lo;
    }
    return sum;
}

// csmat.rs:1782-1829  raw::convert_mat_storage : counting-sort transpose
// (CSR<->CSC).  `inner` = inner dimension of the input.
template <class I, class P>
void convert_mat_storage(size_t outer, size_t inner, const P* indptr,
                         const I* indices, const double* data, P* o_indptr,
                         I* o_indices, double* o_data) {
    const P base = indptr[0];
    for (size_t i = 0; i <= inner; ++i) o_indptr[i] = 0;
    for (size_t o = 0; o < outer; ++o)
        for (size_t k = indptr[o] - base, e = indptr[o + 1] - base; k < e; ++k)
            o_indptr[indices[k]] += 1;
    P cumsum = 0;
    for (size_t i = 0; i <= inner; ++i) {
        const P tmp = o_indptr[i];
        o_indptr[i] = cumsum;
        cumsum += tmp;
    }
    for (size_t o = 0; o < outer; ++o)
        for (size_t k = indptr[o] - base, e = indptr[o + 1] - base; k < e; ++k) {
            const size_t dest = o_indptr[indices[k]];
            o_data[dest] = data[k];
            o_indices[dest] = (I)o;
            o_indptr[indices[k]] += 1;
        }
    P last = 0;
    for (size_t i = 0; i <= inner; ++i) std::swap(o_indptr[i], last);
}

// triplet_iter.rs:127-224  TriMatIter::into_cs (CSR arm): sort the (row, col, value)
// records by (row, col), add each duplicate into the slot of its first occurrence, build
// indptr while walking.  The reference sorts with sort_unstable_by_key, so the order in
// which duplicates are added is unspecified there; std::stable_sort (insertion order) is
// one of its valid outcomes and is what the device path reproduces.
template <class I, class P>
size_t triplets_to_csr(size_t rows, size_t n, const I* ri, const I* ci, const double* v,
                       P* indptr, I* indices, double* data) {
    struct Rec { I r, c; double v; };
    std::vector<Rec> rc(n);
    for (size_t k = 0; k < n; ++k) rc[k] = Rec{ri[k], ci[k], v[k]};
    std::stable_sort(rc.begin(), rc.end(), [](const Rec& a, const Rec& b) {
        return a.r != b.r ? a.r < b.r : a.c < b.c;
    });
    size_t slot = 0, cur_outer = 0;
    for (size_t o = 0; o <= rows; ++o) indptr[o] = 0;
    for (size_t rec = 0; rec < n; ++rec) {
        if (rec > 0) {
            if (rc[rec - 1].r == rc[rec].r && rc[rec - 1].c == rc[rec].c) {
                rc[slot].v = rc[slot].v + rc[rec].v;  // duplicate: add into the current slot
            } else {
                slot += 1;
                rc[slot] = rc[rec];
            }
        }
        const size_t new_outer = rc[rec].r;
        while (new_outer > cur_outer) {
            indptr[cur_outer + 1] = (P)slot;
            cur_outer += 1;
        }
    }
    if (n > 0) slot += 1;
    while (rows > cur_outer) {
        indptr[cur_outer + 1] = (P)slot;
        cur_outer += 1;
    }
    for (size_t k = 0; k < slot; ++k) {
        indices[k] = rc[k].c;
        data[k] = rc[k].v;
    }
    return slot;
}

// ---------------------------------------------------------------------------
// smmp.rs:81-131  symbolic : pattern of C = A*B for a chunk of A rows.
// `seen` has b_cols entries.  Appends to c_indices; c_indptr has a_rows+1
// entries and is zero-based for the chunk.
template <class I, class P>
void symbolic(size_t a_rows, const P* a_indptr, const I* a_indices,
              const P* b_indptr, const I* b_indices, P* c_indptr,
              std::vector<I>& c_indices, unsigned char* seen, size_t b_cols) {
    const P a_base = a_indptr[0], b_base = b_indptr[0];
    c_indices.clear();
    std::memset(seen, 0, b_cols);
    c_indptr[0] = 0;
    for (size_t a_row = 0; a_row < a_rows; ++a_row) {
        size_t length = 0;
        for (size_t ka = a_indptr[a_row] - a_base, ea = a_indptr[a_row + 1] - a_base;
             ka < ea; ++ka) {
            const size_t b_row = a_indices[ka];
            for (size_t kb = b_indptr[b_row] - b_base, eb = b_indptr[b_row + 1] - b_base;
                 kb < eb; ++kb) {
                const size_t b_col = b_indices[kb];
                if (!seen[b_col]) {
                    seen[b_col] = 1;
                    c_indices.push_back((I)b_col);
                    ++length;
                }
            }
        }
        c_indptr[a_row + 1] = c_indptr[a_row] + (P)length;
        const size_t c_start = c_indptr[a_row];
        std::sort(c_indices.begin() + c_start, c_indices.begin() + c_start + length);
        for (size_t q = c_start; q < c_start + length; ++q) seen[c_indices[q]] = 0;
    }
}

// smmp.rs:151-189  numeric : values of C for a chunk of rows, dense accumulator
// `tmp` (b_cols entries), gather in C's (sorted) column order, reset to zero.
// c_indptr is the chunk's slice of the global indptr (non-zero-based allowed);
// c_indices / c_data point at the chunk's first element.
template <class I, class P>
void numeric(size_t a_rows, const P* a_indptr, const I* a_indices,
             const double* a_data, const P* b_indptr, const I* b_indices,
             const double* b_data, const P* c_indptr, const I* c_indices,
             double* c_data, double* tmp, size_t b_cols) {
    const P a_base = a_indptr[0], b_base = b_indptr[0], c_base = c_indptr[0];
    for (size_t j = 0; j < b_cols; ++j) tmp[j] = 0.0;
    for (size_t r = 0; r < a_rows; ++r) {
        for (size_t ka = a_indptr[r] - a_base, ea = a_indptr[r + 1] - a_base; ka < ea; ++ka) {
            const size_t b_row = a_indices[ka];
            const double a_val = a_data[ka];
            for (size_t kb = b_indptr[b_row] - b_base, eb = b_indptr[b_row + 1] - b_base;
                 kb < eb; ++kb) {
                const double prod = a_val * b_data[kb];
                tmp[b_indices[kb]] = tmp[b_indices[kb]] + prod;
            }
        }
        for (size_t q = c_indptr[r] - c_base, e = c_indptr[r + 1] - c_base; q < e; ++q) {
            c_data[q] = tmp[c_indices[q]];
            tmp[c_indices[q]] = 0.0;
        }
    }
}

// smmp.rs:210-227  thread-count rule of mul_csr_csr (ThreadingStrategy::Automatic
// when requested == 0, Fixed(n) otherwise), clamped by rows.max(1).
size_t smmp_nb_threads(size_t a_rows, size_t a_nnz, size_t b_nnz, size_t requested,
                       size_t nb_cpus) {
    size_t want;
    if (requested > 0) {
        want = requested;
    } else {
        const size_t ideal_chunk_size = 8128;
        const size_t wanted_threads = (a_nnz + b_nnz) / ideal_chunk_size;
        want = std::min(std::max<size_t>(1, wanted_threads), nb_cpus);
    }
    return std::min(std::max<size_t>(a_rows, 1), want);
}

// smmp.rs:256-416  mul_csr_csr_with_workspace : chunked two-phase driver.
// Phase 1: nb_threads equal-row chunks (chunk_size = indptr.len()/nb_threads,
// smmp.rs:277-296), symbolic per chunk, serial concat + prefix sum (320-331).
// Phase 2: rows split into chunks of ~nnzC/nb_threads (332-372), numeric per
// chunk.  OpenMP stands in for rayon; results are thread-count independent.
// Caller frees nothing: outputs are returned through a handle (see extern C).
template <class I, class P>
struct SpgemmResult {
    std::vector<P> indptr;
    std::vector<I> indices;
    std::vector<double> data;
};

template <class I, class P>
SpgemmResult<I, P>* mul_csr_csr(size_t a_rows, size_t a_cols, size_t b_cols,
                                const P* a_indptr, const I* a_indices,
                                const double* a_data, const P* b_indptr,
                                const I* b_indices, const double* b_data,
                                size_t requested_threads) {
    (void)a_cols;
    const size_t a_nnz = a_indptr[a_rows] - a_indptr[0];
    size_t nb_cpus = 1;
#ifdef _OPENMP
    nb_cpus = (size_t)omp_get_num_procs();
#endif
    const size_t b_rows_nnz = b_indptr[a_cols] - b_indptr[0];
    const size_t nb_threads =
        smmp_nb_threads(a_rows, a_nnz, b_rows_nnz, requested_threads, nb_cpus);

    auto* res = new SpgemmResult<I, P>();
    // ---- phase 1: symbolic over equal-row chunks (smmp.rs:277-319)
    const size_t chunk_size = (a_rows + 1) / nb_threads;
    std::vector<size_t> starts(nb_threads), stops(nb_threads);
    for (size_t c = 0; c < nb_threads; ++c) {
        starts[c] = c == 0 ? 0 : c * chunk_size;
        stops[c] = (c + 1 < nb_threads) ? (c + 1) * chunk_size : a_rows;
    }
    std::vector<std::vector<P>> ip_chunks(nb_threads);
    std::vector<std::vector<I>> ind_chunks(nb_threads);
#pragma omp parallel for schedule(static, 1) num_threads((int)nb_threads)
    for (long c = 0; c < (long)nb_threads; ++c) {
        std::vector<unsigned char> seen(b_cols);
        const size_t n = stops[c] - starts[c];
        ip_chunks[c].assign(n + 1, 0);
        symbolic<I, P>(n, a_indptr + starts[c],
                       a_indices + (a_indptr[starts[c]] - a_indptr[0]), b_indptr, b_indices,
                       ip_chunks[c].data(), ind_chunks[c], seen.data(), b_cols);
    }
    // serial concat + prefix sum (smmp.rs:320-331)
    size_t total = 0;
    for (auto& v : ind_chunks) total += v.size();
    res->indices.reserve(total);
    for (auto& v : ind_chunks) res->indices.insert(res->indices.end(), v.begin(), v.end());
    res->indptr.reserve(a_rows + 1);
    res->indptr.push_back(0);
    for (auto& ip : ip_chunks)
        for (size_t r = 0; r + 1 < ip.size(); ++r)
            res->indptr.push_back((P)(ip[r + 1] - ip[r]) + res->indptr.back());
    res->data.assign(res->indices.size(), 0.0);

    // ---- phase 2: numeric over ~equal-nnz row chunks (smmp.rs:332-404).
    // The reference's split rule: a new chunk starts at row-1 whenever the nnz
    // since the last split exceeds nnzC/nb_threads.
    const size_t nnz_chunk = res->indices.size() / nb_threads;
    std::vector<size_t> split_rows;  // chunk c covers rows [split_rows[c], split_rows[c+1])
    split_rows.push_back(0);
    {
        size_t split_nnz = 0;
        for (size_t row = 0; row < res->indptr.size(); ++row) {
            const size_t nnz = res->indptr[row];
            if (nnz - split_nnz > nnz_chunk && row > 0) {
                split_rows.push_back(row - 1);
                split_nnz = nnz;  // as in the reference (smmp.rs:364)
            }
        }
        split_rows.push_back(a_rows);
    }
    const long n_chunks = (long)split_rows.size() - 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads((int)nb_threads)
    for (long c = 0; c < n_chunks; ++c) {
        const size_t r0 = split_rows[c], r1 = split_rows[c + 1];
        if (r1 <= r0) continue;
        std::vector<double> tmp(b_cols);
        const P* cip = res->indptr.data() + r0;
        numeric<I, P>(r1 - r0, a_indptr + r0, a_indices + (a_indptr[r0] - a_indptr[0]),
                      a_data + (a_indptr[r0] - a_indptr[0]), b_indptr, b_indices, b_data, cip,
                      res->indices.data() + cip[0], res->data.data() + cip[0], tmp.data(),
                      b_cols);
    }
    return res;
}

// ---------------------------------------------------------------------------
// linalg/bicgstab.rs:95-300  BiCGSTAB (the iterative caller of SpMV).
//
// The reference keeps every vector as a CsVec and builds the iteration from sparse ops:
// `&A * &v` (vec.rs:1104-1131 -> prod.rs:162-184 for CSR, csmat_mul_csmat for CSC),
// csvec_binop (binop.rs:442-470: a missing entry takes part as an explicit 0.0),
// `map` (vec.rs:989-997), `dot` (vec.rs:846-881: sequential mul_acc over the matching
// entries) and `squared_l2_norm` (vec.rs:907-913: sequential sum of x*x).  On vectors whose
// stored pattern is full that is exactly dense arithmetic in index order, which is what is
// restated here (a dropped exact zero contributes +0.0 to every sum, so the value is the
// same; only NaN/Inf-times-zero and the sign of zero could differ).  Both storages give the
// same summation order for A*v: ascending column within each row.
struct Bicgstab {
    size_t n = 0;
    std::function<void(const double*, double*)> matvec;  // y = A x (y zeroed first)
    std::vector<double> b, x, r, rhat, p;
    double rho = 0, err = 0, soft_restart_threshold = 0.1;
    size_t iteration_count = 0, soft_restart_count = 0, hard_restart_count = 0;

    static double dot(const std::vector<double>& a, const std::vector<double>& c) {
        double sum = 0.0;
        for (size_t i = 0; i < a.size(); ++i) {
            const double prod = a[i] * c[i];
            sum = sum + prod;
        }
        return sum;
    }
    // bicgstab.rs:120-146  new
    void init(const double* x0, const double* b_) {
        b.assign(b_, b_ + n);
        x.assign(x0, x0 + n);
        std::vector<double> ax(n, 0.0);
        matvec(x.data(), ax.data());
        r.resize(n);
        for (size_t i = 0; i < n; ++i) r[i] = b[i] - ax[i];
        rhat = r;
        p = r;
        err = std::sqrt(dot(r, r));
        rho = err * err;
    }
    // bicgstab.rs:177-184
    void soft_restart() {
        soft_restart_count += 1;
        rhat = r;
        rho = err * err;
        p = r;
    }
    // bicgstab.rs:186-196
    void hard_restart() {
        hard_restart_count += 1;
        std::vector<double> ax(n, 0.0);
        matvec(x.data(), ax.data());
        for (size_t i = 0; i < n; ++i) r[i] = b[i] - ax[i];
        err = std::sqrt(dot(r, r));
        soft_restart();
        soft_restart_count -= 1;
    }
    // bicgstab.rs:198-234
    double step() {
        iteration_count += 1;
        std::vector<double> v(n, 0.0), s(n), t(n, 0.0), h(n);
        matvec(p.data(), v.data());
        const double alpha = rho / dot(rhat, v);
        for (size_t i = 0; i < n; ++i) {
            const double pa = p[i] * alpha;
            h[i] = x[i] + pa;
        }
        for (size_t i = 0; i < n; ++i) {
            const double va = v[i] * alpha;
            s[i] = r[i] - va;
        }
        matvec(s.data(), t.data());
        const double omega = dot(t, s) / dot(t, t);
        for (size_t i = 0; i < n; ++i) {
            const double os = omega * s[i];
            x[i] = h[i] + os;
        }
        for (size_t i = 0; i < n; ++i) {
            const double to = t[i] * omega;
            r[i] = s[i] - to;
        }
        err = std::sqrt(dot(r, r));
        const double rho_prev = rho;
        rho = dot(rhat, r);
        if (std::fabs(rho) / (err * err) < soft_restart_threshold) {
            soft_restart();
        } else {
            const double beta = (rho / rho_prev) * (alpha / omega);
            for (size_t i = 0; i < n; ++i) {
                const double vo = v[i] * omega;
                const double d = p[i] - vo;
                const double db = d * beta;
                p[i] = r[i] + db;
            }
        }
        return err;
    }
    // bicgstab.rs:151-175: 1 = Ok, 0 = Err (the solver state is returned either way)
    int solve(double tol, size_t max_iter) {
        for (size_t it = 0; it < max_iter; ++it) {
            step();
            if (err < tol) {
                hard_restart();
                if (err < tol) return 1;
            }
        }
        return 0;
    }
};

template <class I, class P>
Bicgstab* bicgstab_new(size_t rows, const P* ip, const I* ind, const double* d, const double* x0,
                       const double* b) {
    auto* s = new Bicgstab;
    s->n = rows;
    s->matvec = [=](const double* x, double* y) {
        mul_acc_mat_vec_csr<I, P>(rows, ip, ind, d, x, y);
    };
    s->init(x0, b);
    return s;
}

}  // namespace

// ---------------------------------------------------------------------------
// C exports for ctypes.  Suffix _IP: index bytes / indptr bytes (44, 88, 48).
#define ORACLE_EXPORTS(SUF, I, P)                                                              \
    extern "C" void oracle_mul_acc_mat_vec_csr_##SUF(size_t rows, const P* ip, const I* ind,    \
                                                     const double* d, const double* x,          \
                                                     double* y) {                               \
        mul_acc_mat_vec_csr<I, P>(rows, ip, ind, d, x, y);                                      \
    }                                                                                           \
    extern "C" void oracle_mul_acc_mat_vec_csc_##SUF(size_t cols, const P* ip, const I* ind,    \
                                                     const double* d, const double* x,          \
                                                     double* y) {                               \
        mul_acc_mat_vec_csc<I, P>(cols, ip, ind, d, x, y);                                      \
    }                                                                                           \
    extern "C" void oracle_csr_mulacc_dense_colmaj_##SUF(                                       \
        size_t rows, size_t k, const P* ip, const I* ind, const double* d, const double* rhs,   \
        ptrdiff_t rrs, ptrdiff_t rcs, double* out, ptrdiff_t ors, ptrdiff_t ocs) {              \
        csr_mulacc_dense_colmaj<I, P>(rows, k, ip, ind, d, rhs, rrs, rcs, out, ors, ocs);       \
    }                                                                                           \
    extern "C" void oracle_csr_mulacc_dense_rowmaj_##SUF(                                       \
        size_t rows, size_t k, const P* ip, const I* ind, const double* d, const double* rhs,   \
        ptrdiff_t rrs, ptrdiff_t rcs, double* out, ptrdiff_t ors, ptrdiff_t ocs) {              \
        csr_mulacc_dense_rowmaj<I, P>(rows, k, ip, ind, d, rhs, rrs, rcs, out, ors, ocs);       \
    }                                                                                           \
    extern "C" void oracle_csc_mulacc_dense_colmaj_##SUF(                                       \
        size_t cols, size_t k, const P* ip, const I* ind, const double* d, const double* rhs,   \
        ptrdiff_t rrs, ptrdiff_t rcs, double* out, ptrdiff_t ors, ptrdiff_t ocs) {              \
        csc_mulacc_dense_colmaj<I, P>(cols, k, ip, ind, d, rhs, rrs, rcs, out, ors, ocs);       \
    }                                                                                           \
    extern "C" void oracle_csc_mulacc_dense_rowmaj_##SUF(                                       \
        size_t cols, size_t k, const P* ip, const I* ind, const double* d, const double* rhs,   \
        ptrdiff_t rrs, ptrdiff_t rcs, double* out, ptrdiff_t ors, ptrdiff_t ocs) {              \
        csc_mulacc_dense_rowmaj<I, P>(cols, k, ip, ind, d, rhs, rrs, rcs, out, ors, ocs);       \
    }                                                                                           \
    extern "C" size_t oracle_csr_mul_csvec_##SUF(size_t rows, const P* ip, const I* ind,        \
                                                 const double* d, size_t vn, const I* vi,       \
                                                 const double* vd, I* oi, double* od) {         \
        return csr_mul_csvec<I, P>(rows, ip, ind, d, vn, vi, vd, oi, od);                       \
    }                                                                                           \
    extern "C" double oracle_csvec_dot_by_binary_search_##SUF(size_t n1, const I* i1,           \
                                                              const double* d1, size_t n2,      \
                                                              const I* i2, const double* d2) {  \
        (void)sizeof(P);                                                                        \
        return csvec_dot_by_binary_search<I>(n1, i1, d1, n2, i2, d2);                           \
    }                                                                                           \
    extern "C" void oracle_convert_mat_storage_##SUF(size_t outer, size_t inner, const P* ip,   \
                                                     const I* ind, const double* d, P* oip,     \
                                                     I* oind, double* od) {                     \
        convert_mat_storage<I, P>(outer, inner, ip, ind, d, oip, oind, od);                     \
    }                                                                                           \
    extern "C" size_t oracle_triplets_to_csr_##SUF(size_t rows, size_t n, const I* ri,            \
                                                   const I* ci, const double* v, P* ip, I* ind,  \
                                                   double* d) {                                  \
        return triplets_to_csr<I, P>(rows, n, ri, ci, v, ip, ind, d);                            \
    }                                                                                           \
    extern "C" void* oracle_bicgstab_new_##SUF(size_t rows, const P* ip, const I* ind,           \
                                               const double* d, const double* x0,               \
                                               const double* b) {                               \
        return bicgstab_new<I, P>(rows, ip, ind, d, x0, b);                                     \
    }                                                                                           \
    extern "C" void* oracle_mul_csr_csr_##SUF(size_t ar, size_t ac, size_t bc, const P* aip,    \
                                              const I* aind, const double* ad, const P* bip,    \
                                              const I* bind, const double* bd,                  \
                                              size_t threads) {                                 \
        return mul_csr_csr<I, P>(ar, ac, bc, aip, aind, ad, bip, bind, bd, threads);            \
    }                                                                                           \
    extern "C" size_t oracle_spgemm_nnz_##SUF(void* h) {                                        \
        return ((SpgemmResult<I, P>*)h)->indices.size();                                        \
    }                                                                                           \
    extern "C" void oracle_spgemm_fetch_##SUF(void* h, P* ip, I* ind, double* d) {              \
        auto* r = (SpgemmResult<I, P>*)h;                                                       \
        std::copy(r->indptr.begin(), r->indptr.end(), ip);                                      \
        std::copy(r->indices.begin(), r->indices.end(), ind);                                   \
        std::copy(r->data.begin(), r->data.end(), d);                                           \
    }                                                                                           \
    extern "C" void oracle_spgemm_free_##SUF(void* h) { delete (SpgemmResult<I, P>*)h; }        \
    /* symbolic + numeric called separately, as smmp.rs:423-465 does */                        \
    extern "C" size_t oracle_symbolic_##SUF(size_t ar, size_t bc, const P* aip, const I* aind,  \
                                            const P* bip, const I* bind, P* cip, I* cind,       \
                                            size_t cind_cap) {                                  \
        std::vector<I> ci;                                                                      \
        std::vector<unsigned char> seen(bc);                                                    \
        symbolic<I, P>(ar, aip, aind, bip, bind, cip, ci, seen.data(), bc);                     \
        if (ci.size() <= cind_cap) std::copy(ci.begin(), ci.end(), cind);                       \
        return ci.size();                                                                       \
    }                                                                                           \
    extern "C" void oracle_numeric_##SUF(size_t ar, size_t bc, const P* aip, const I* aind,     \
                                         const double* ad, const P* bip, const I* bind,         \
                                         const double* bd, const P* cip, const I* cind,         \
                                         double* cd) {                                          \
        std::vector<double> tmp(bc);                                                            \
        numeric<I, P>(ar, aip, aind, ad, bip, bind, bd, cip, cind, cd, tmp.data(), bc);         \
    }

ORACLE_EXPORTS(44, uint32_t, uint32_t)
ORACLE_EXPORTS(88, uint64_t, uint64_t)
ORACLE_EXPORTS(48, uint32_t, uint64_t)

// All-cores row-chunked SpMV -- an EXTENSION that is NOT in the reference
// (sprs SpMV is single-threaded).  Contiguous row blocks
// (slice_outer semantics, slicing.rs:65-89) under OpenMP static scheduling.
// Used only by bench.py as an additional, clearly-labelled CPU number.
extern "C" void oracle_ext_spmv_csr_omp_44(size_t rows, const uint32_t* ip, const uint32_t* ind,
                                           const double* d, const double* x, double* y,
                                           int threads) {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long r = 0; r < (long)rows; ++r) {
        double tv = y[r];
        for (size_t k = ip[r] - ip[0], e = ip[r + 1] - ip[0]; k < e; ++k) {
            const double prod = d[k] * x[ind[k]];
            tv = tv + prod;
        }
        y[r] = tv;
    }
}

// BiCGSTAB handle API (the matrix arrays passed to oracle_bicgstab_new_* are borrowed).
extern "C" double oracle_bicgstab_step(void* h) { return ((Bicgstab*)h)->step(); }
extern "C" void oracle_bicgstab_soft_restart(void* h) { ((Bicgstab*)h)->soft_restart(); }
extern "C" void oracle_bicgstab_hard_restart(void* h) { ((Bicgstab*)h)->hard_restart(); }
extern "C" int oracle_bicgstab_solve(void* h, double tol, size_t max_iter) {
    return ((Bicgstab*)h)->solve(tol, max_iter);
}
extern "C" void oracle_bicgstab_set_threshold(void* h, double t) {
    ((Bicgstab*)h)->soft_restart_threshold = t;
}
// which: 0 x, 1 r, 2 rhat, 3 p, 4 b
extern "C" void oracle_bicgstab_get(void* h, int which, double* out) {
    Bicgstab* s = (Bicgstab*)h;
    const std::vector<double>* v[5] = {&s->x, &s->r, &s->rhat, &s->p, &s->b};
    std::memcpy(out, v[which]->data(), s->n * sizeof(double));
}
// counts = {iteration, soft restarts, hard restarts}; scalars = {err, rho, threshold}
extern "C" void oracle_bicgstab_stats(void* h, size_t* counts, double* scalars) {
    Bicgstab* s = (Bicgstab*)h;
    counts[0] = s->iteration_count;
    counts[1] = s->soft_restart_count;
    counts[2] = s->hard_restart_count;
    scalars[0] = s->err;
    scalars[1] = s->rho;
    scalars[2] = s->soft_restart_threshold;
}
extern "C" void oracle_bicgstab_free(void* h) { delete (Bicgstab*)h; }

extern "C" int oracle_num_procs(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
