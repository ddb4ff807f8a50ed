/*
 * ORACLE -- test infrastructure only.  Never linked, imported or executed by the product
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * CPU restatement of the arithmetic the reference's hot path bottoms out in.  The reference
 * (spcl/arrow-matrix) has no native code of its own: every product is SciPy
 * `csr_matrix @ ndarray` (arrow/arrow_slim_mpi.py:109-111, 125-127, 142-144;
 * arrow/arrow_mpi.py:198-219, 289-291), which dispatches to the third-party kernel
 * scipy.sparse._sparsetools.csr_matvecs (SciPy >= 1.12, container has 1.18.1; not vendored
 * under /root/reference).  Its published algorithm, restated:
 *
 *     for each row i:  for each stored entry jj of row i (in storage order):
 *         y[i, :] += a[jj] * x[col[jj], :]          (an axpy over the k feature columns)
 *
 * with row-major X and Y and sequential fp32 accumulation.  Checked bit-for-bit against
 * SciPy itself in tests/test_oracle.py.
 */
#include <stdint.h>
#include <stddef.h>

/* Y (n_row x k, row-major) += A (CSR) * X (n_col x k, row-major). */
void oracle_csr_matvecs_f32_i32(int64_t n_row, int64_t k, const int32_t *indptr,
                                const int32_t *indices, const float *data,
                                const float *restrict x, float *restrict y)
{
    for (int64_t i = 0; i < n_row; ++i) {
        float *restrict yi = y + (size_t)i * (size_t)k;
        for (int32_t jj = indptr[i]; jj < indptr[i + 1]; ++jj) {
            const float a = data[jj];
            const float *restrict xj = x + (size_t)indices[jj] * (size_t)k;
            for (int64_t c = 0; c < k; ++c)
                yi[c] += a * xj[c];
        }
    }
}

void oracle_csr_matvecs_f32_i64(int64_t n_row, int64_t k, const int64_t *indptr,
                                const int64_t *indices, const float *data,
                                const float *restrict x, float *restrict y)
{
    for (int64_t i = 0; i < n_row; ++i) {
        float *restrict yi = y + (size_t)i * (size_t)k;
        for (int64_t jj = indptr[i]; jj < indptr[i + 1]; ++jj) {
            const float a = data[jj];
            const float *restrict xj = x + (size_t)indices[jj] * (size_t)k;
            for (int64_t c = 0; c < k; ++c)
                yi[c] += a * xj[c];
        }
    }
}

/* Row-range variant so a host thread pool (bench.py --impl reference) can split rows the way
 * the reference splits block-rows over MPI ranks. */
void oracle_csr_matvecs_f32_i32_rows(int64_t row_begin, int64_t row_end, int64_t k,
                                     const int32_t *indptr, const int32_t *indices,
                                     const float *data, const float *restrict x, float *restrict y)
{
    for (int64_t i = row_begin; i < row_end; ++i) {
        float *restrict yi = y + (size_t)i * (size_t)k;
        for (int32_t jj = indptr[i]; jj < indptr[i + 1]; ++jj) {
            const float a = data[jj];
            const float *restrict xj = x + (size_t)indices[jj] * (size_t)k;
            for (int64_t c = 0; c < k; ++c)
                yi[c] += a * xj[c];
        }
    }
}

/* Exchange steps of the reference (numpy fancy indexing there):
 *   forward   C_i[recv_perm] = recvbuf            arrow/arrow_dec_mpi.py:544
 *   backward  C_i[recv_perm] += recvbuf           arrow/arrow_dec_mpi.py:437
 * in global form: dst[r] = src[map[r]]  /  dst[map[r]] += src[r]; map[r] < 0 or >= limit is the
 * sentinel and is skipped. */
void oracle_gather_rows_f32(int64_t n_dst, int64_t k, const int64_t *map, int64_t src_rows,
                            const float *src, float *dst)
{
    for (int64_t r = 0; r < n_dst; ++r) {
        int64_t s = map[r];
        if (s < 0 || s >= src_rows) continue;
        for (int64_t c = 0; c < k; ++c) dst[(size_t)r * k + c] = src[(size_t)s * k + c];
    }
}

void oracle_scatter_add_rows_f32(int64_t n_src, int64_t k, const int64_t *map, int64_t dst_rows,
                                 const float *src, float *dst)
{
    for (int64_t r = 0; r < n_src; ++r) {
        int64_t d = map[r];
        if (d < 0 || d >= dst_rows) continue;
        for (int64_t c = 0; c < k; ++c) dst[(size_t)d * k + c] += src[(size_t)r * k + c];
    }
}

/* Row-range forms of the two exchanges so host threads can split them like MPI ranks split tiles. */
void oracle_gather_rows_f32_range(int64_t r_begin, int64_t r_end, int64_t k, const int64_t *map,
                                  int64_t src_rows, const float *src, float *dst)
{
    for (int64_t r = r_begin; r < r_end; ++r) {
        int64_t s = map[r];
        if (s < 0 || s >= src_rows) continue;
        const float *sp = src + (size_t)s * k;
        float *dp = dst + (size_t)r * k;
        for (int64_t c = 0; c < k; ++c) dp[c] = sp[c];
    }
}

void oracle_scatter_add_rows_f32_range(int64_t r_begin, int64_t r_end, int64_t k, const int64_t *map,
                                       int64_t dst_rows, const float *src, float *dst)
{
    for (int64_t r = r_begin; r < r_end; ++r) {
        int64_t d = map[r];
        if (d < 0 || d >= dst_rows) continue;
        const float *sp = src + (size_t)r * k;
        float *dp = dst + (size_t)d * k;
        for (int64_t c = 0; c < k; ++c) dp[c] += sp[c];
    }
}

void oracle_zero_rows_f32(int64_t r_begin, int64_t r_end, int64_t k, float *dst)
{
    for (size_t i = (size_t)r_begin * k; i < (size_t)r_end * k; ++i) dst[i] = 0.0f;
}
