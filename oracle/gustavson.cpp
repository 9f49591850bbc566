// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
//
// Host oracle for C = A*B (A^2 and A*A^T) in fp64: OpenMP row-wise Gustavson.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The CUDA engine never calls it and has no CPU fallback.
//
// Parity note: the reference (/root/reference) has NO CPU SpGEMM and NO tests or golden
// vectors, so this oracle restates the *math* the reference's GPU
// kernels implement, following these reference sites:
//   - numeric accumulation order  spgemm.cu:593-661  (pairs visited in ascending tile-k,
//     bits inside a tile pair in ascending k via __ffs, one DFMA per product)  ==>
//     C[i,j] = fma(a_ik, b_kj, C[i,j]) for k ascending, starting from +0.0;
//   - structural result           spgemm.cu:499-550  (C's structure is the boolean
//     product of the masks; numerically cancelled entries are KEPT)  ==>  symbolic
//     Gustavson, no zero dropping;
//   - flop                        spgemm.cu:1068-1079 (flop = sum_{a_ik} nnz(B_k,:)).
// Pins: config-1 known answers in SURVEY.md section 8c (scipy, survey time) and the
// reference's own sm_100 rebuild run on a B200 (tests/golden/, see tests/golden/README.md).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

extern "C" {

int oracle_max_threads() { return omp_get_max_threads(); }
// launchers such as torchrun export OMP_NUM_THREADS=1: a caller that wants all host cores says so
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// COO (any order, no duplicates) -> CSR with columns ascending inside each row.
// Returns 0, or -1 if a coordinate is out of range, -2 if a duplicate (i,j) exists.
int oracle_coo_to_csr(int rows, int cols, int64_t nnz, const int* I, const int* J, const double* V,
                      int64_t* rowptr, int* colidx, double* vals)
{
    std::fill(rowptr, rowptr + rows + 1, (int64_t)0);
    for (int64_t e = 0; e < nnz; ++e) {
        if (I[e] < 0 || I[e] >= rows || J[e] < 0 || J[e] >= cols) return -1;
        ++rowptr[I[e] + 1];
    }
    for (int r = 0; r < rows; ++r) rowptr[r + 1] += rowptr[r];
    std::vector<int64_t> cur(rowptr, rowptr + rows);
    for (int64_t e = 0; e < nnz; ++e) {
        int64_t p = cur[I[e]]++;
        colidx[p] = J[e];
        vals[p] = V[e];
    }
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(| : bad)
    for (int r = 0; r < rows; ++r) {
        int64_t b = rowptr[r], e = rowptr[r + 1];
        std::vector<std::pair<int, double>> tmp((size_t)(e - b));
        for (int64_t p = b; p < e; ++p) tmp[(size_t)(p - b)] = {colidx[p], vals[p]};
        std::sort(tmp.begin(), tmp.end(), [](auto const& x, auto const& y) { return x.first < y.first; });
        for (int64_t p = b; p < e; ++p) {
            colidx[p] = tmp[(size_t)(p - b)].first;
            vals[p] = tmp[(size_t)(p - b)].second;
            if (p > b && colidx[p] == colidx[p - 1]) bad = 1;
        }
    }
    return bad ? -2 : 0;
}

// CSR transpose (columns ascending inside each output row because input rows are visited ascending).
void oracle_csr_transpose(int rows, int cols, const int64_t* Ap, const int* Aj, const double* Ax,
                          int64_t* Tp, int* Tj, double* Tx)
{
    std::fill(Tp, Tp + cols + 1, (int64_t)0);
    for (int64_t p = 0; p < Ap[rows]; ++p) ++Tp[Aj[p] + 1];
    for (int c = 0; c < cols; ++c) Tp[c + 1] += Tp[c];
    std::vector<int64_t> cur(Tp, Tp + cols);
    for (int r = 0; r < rows; ++r)
        for (int64_t p = Ap[r]; p < Ap[r + 1]; ++p) {
            int64_t q = cur[Aj[p]]++;
            Tj[q] = r;
            Tx[q] = Ax[p];
        }
}

// flop = sum over nonzeros a_ik of nnz(B row k)   (reference: spgemm.cu:1074-1078)
uint64_t oracle_flop(int m, const int64_t* Ap, const int* Aj, const int64_t* Bp)
{
    uint64_t flop = 0;
#pragma omp parallel for schedule(static) reduction(+ : flop)
    for (int i = 0; i < m; ++i)
        for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) flop += (uint64_t)(Bp[Aj[p] + 1] - Bp[Aj[p]]);
    return flop;
}

// Symbolic pass: Cp[i+1]-Cp[i] = number of structurally reachable columns of row i.
// Returns nnz(C).
int64_t oracle_spgemm_symbolic(int m, int n, const int64_t* Ap, const int* Aj,
                               const int64_t* Bp, const int* Bj, int64_t* Cp)
{
    Cp[0] = 0;
#pragma omp parallel
    {
        std::vector<int> mark((size_t)n, -1);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < m; ++i) {
            int64_t cnt = 0;
            for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
                int k = Aj[p];
                for (int64_t q = Bp[k]; q < Bp[k + 1]; ++q) {
                    int j = Bj[q];
                    if (mark[(size_t)j] != i) { mark[(size_t)j] = i; ++cnt; }
                }
            }
            Cp[i + 1] = cnt;
        }
    }
    for (int i = 0; i < m; ++i) Cp[i + 1] += Cp[i];
    return Cp[m];
}

// Numeric pass into preallocated Cj/Cx (sizes from the symbolic pass).
// Row i: for k ascending over A's row, for each b_kj: acc[j] = fma(a_ik, b_kj, acc[j]).
void oracle_spgemm_numeric(int m, int n, const int64_t* Ap, const int* Aj, const double* Ax,
                           const int64_t* Bp, const int* Bj, const double* Bx,
                           const int64_t* Cp, int* Cj, double* Cx)
{
#pragma omp parallel
    {
        std::vector<int> mark((size_t)n, -1);
        std::vector<double> acc((size_t)n, 0.0);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < m; ++i) {
            int64_t base = Cp[i], cnt = 0;
            for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
                int k = Aj[p];
                double a = Ax[p];
                for (int64_t q = Bp[k]; q < Bp[k + 1]; ++q) {
                    int j = Bj[q];
                    if (mark[(size_t)j] != i) {
                        mark[(size_t)j] = i;
                        acc[(size_t)j] = 0.0;
                        Cj[base + cnt++] = j;
                    }
                    acc[(size_t)j] = std::fma(a, Bx[q], acc[(size_t)j]);
                }
            }
            std::sort(Cj + base, Cj + base + cnt);
            for (int64_t t = 0; t < cnt; ++t) Cx[base + t] = acc[(size_t)Cj[base + t]];
        }
    }
}

// fp32 instantiation of the numeric pass (the reference's kernels are templates over ValueType,
// spgemm.cu:593,727-728): same order, single-precision fma.
void oracle_spgemm_numeric_f32(int m, int n, const int64_t* Ap, const int* Aj, const float* Ax,
                               const int64_t* Bp, const int* Bj, const float* Bx,
                               const int64_t* Cp, int* Cj, float* Cx)
{
#pragma omp parallel
    {
        std::vector<int> mark((size_t)n, -1);
        std::vector<float> acc((size_t)n, 0.0f);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < m; ++i) {
            int64_t base = Cp[i], cnt = 0;
            for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
                int k = Aj[p];
                float a = Ax[p];
                for (int64_t q = Bp[k]; q < Bp[k + 1]; ++q) {
                    int j = Bj[q];
                    if (mark[(size_t)j] != i) {
                        mark[(size_t)j] = i;
                        acc[(size_t)j] = 0.0f;
                        Cj[base + cnt++] = j;
                    }
                    acc[(size_t)j] = std::fmaf(a, Bx[q], acc[(size_t)j]);
                }
            }
            std::sort(Cj + base, Cj + base + cnt);
            for (int64_t t = 0; t < cnt; ++t) Cx[base + t] = acc[(size_t)Cj[base + t]];
        }
    }
}

}  // extern "C"
