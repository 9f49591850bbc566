// The three SpGEMM steps on tiled operands (one iteration of /root/reference/spgemm.cu:1133-1357).
//
//   step 1  tile-level symbolic: C' = structure(A' * B') with its ordered (A tile, B tile) pair
//           lists.  Replaces tile_spgemm_step1_cuda_spa_kernel / ..._numeric_... (:271-384), the
//           NSPARSE hash path (NSPARSE/spgemm_nsparse_kernel.h) AND the CSC-based pair search
//           pem_spgemm_step2_search_pairs (:387-497): the row-wise expansion that finds a C' tile
//           also enumerates its pairs, so the sorted-list intersection with its binary searches
//           (and B's tile-level CSC, :1033-1062) is not needed at all.
//   step 2  per-tile bitmask symbolic: C tile masks, per-tile nnz, scan, rowColIdx.  Replaces
//           pem_spgemm_step2_compute_CMasksAndOffsets (:499-550) and ..._CrowColIdx (:552-591).
//           Atomic-free.
//   step 3  numeric: one thread per C nonzero, ascending-k fma chain.  Replaces
//           pem_spgemm_step3_accumulate (:593-661).  Atomic-free, C written exactly once.
#include <climits>
#include <chrono>
#include <algorithm>

#include "engine.cuh"

namespace {

// =========================================================================================
// small device helpers
// =========================================================================================
template <int THREADS>
__device__ __forceinline__ unsigned block_sum(unsigned v, unsigned* red /* THREADS/32 + 1 words */)
{
    constexpr int NW = THREADS / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w];
    __syncthreads();
    return s;
}

// In-place exclusive scan of a[0..n) in shared memory by the whole block; `in(i)` gives the input
// for slot i (so the input may be derived from another array).  Returns the total to all threads.
template <int THREADS, class In>
__device__ __forceinline__ unsigned block_scan_exclusive(unsigned* a, int n, In in, unsigned* red)
{
    constexpr int NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + THREADS - 1) / THREADS;
    const int b = min(tid * per, n), e = min(b + per, n);
    unsigned local = 0;
    for (int i = b; i < e; ++i) local += in(i);
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    unsigned woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        unsigned x = red[w];
        if (w < warp) woff += x;
        total += x;
    }
    unsigned run = woff + incl - local;
    for (int i = b; i < e; ++i) {
        unsigned x = in(i);
        a[i] = run;
        run += x;
    }
    __syncthreads();
    return total;
}

// =========================================================================================
// step 1, kernel 1: per tile row of the panel: window [jmin, jmax] of reachable tile columns
// and the number of tile-level products P (warp per tile row)
// =========================================================================================
__global__ void __launch_bounds__(256)
k_row_window(int rb, int re, const int32_t* __restrict__ Arp, const int32_t* __restrict__ Acol,
             const int32_t* __restrict__ Brp, const int32_t* __restrict__ Bcol,
             int2* __restrict__ win, int64_t* __restrict__ scalars)
{
    int row = rb + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (row >= re) return;
    int lane = threadIdx.x & 31;
    int jmin = INT_MAX, jmax = -1;
    unsigned long long P = 0;
    for (int p = Arp[row] + lane; p < Arp[row + 1]; p += 32) {
        int k = Acol[p];
        int bs = Brp[k], be = Brp[k + 1];
        if (be > bs) {
            P += (unsigned)(be - bs);
            jmin = min(jmin, Bcol[bs]);
            jmax = max(jmax, Bcol[be - 1]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
        jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
        P += __shfl_xor_sync(0xffffffffu, P, o);
    }
    if (lane == 0) {
        win[row - rb] = make_int2(jmin, jmax);
        if (jmax >= 0) {
            long long words = ((jmax - (jmin & ~31)) >> 5) + 1;
            atomicMax((long long*)&scalars[SC_MAXWIN], words);
            atomicMax((long long*)&scalars[SC_MAXP], (long long)P);
            atomicAdd((unsigned long long*)&scalars[SC_SUMP], P);
        }
    }
}

// =========================================================================================
// step 1, kernel 2 (count): block per tile row, windowed bitmap accumulator in shared memory.
//   D[row] = number of C' tiles in the row, F[row] = number of (A tile, B tile) pairs kept.
// A pair is dropped (unless keep_empty) when colOcc(A tile) & rowOcc(B tile) == 0, i.e. when the
// 16x16 boolean product of the two tiles is empty.
// =========================================================================================
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_step1_count(int rb, int re, const int32_t* __restrict__ Arp, const int32_t* __restrict__ Acol,
              const uint16_t* __restrict__ AcolOcc, const int32_t* __restrict__ Brp,
              const int32_t* __restrict__ Bcol, const uint16_t* __restrict__ BrowOcc,
              const int2* __restrict__ win, int keep_empty,
              int64_t* __restrict__ D, int64_t* __restrict__ F, int64_t* __restrict__ scalars)
{
    extern __shared__ unsigned sm[];
    __shared__ unsigned red[THREADS / 32 + 1];
    constexpr int NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int row = rb + blockIdx.x; row < re; row += gridDim.x) {
        int2 wnd = win[row - rb];
        if (wnd.y < 0) {
            if (tid == 0) { D[row - rb] = 0; F[row - rb] = 0; }
            continue;
        }
        const int base = wnd.x & ~31;
        const int W = ((wnd.y - base) >> 5) + 1;
        for (int w = tid; w < W; w += THREADS) sm[w] = 0;
        __syncthreads();
        unsigned f = 0;
        const int as = Arp[row], ae = Arp[row + 1];
        for (int p = as + warp; p < ae; p += NW) {
            const int k = Acol[p];
            const unsigned aocc = keep_empty ? 0xFFFFu : AcolOcc[p];
            const int bs = Brp[k], be = Brp[k + 1];
            for (int q = bs + lane; q < be; q += 32) {
                if (aocc & BrowOcc[q]) {
                    int j = Bcol[q] - base;
                    atomicOr(&sm[j >> 5], 1u << (j & 31));
                    ++f;
                }
            }
        }
        __syncthreads();
        unsigned d = 0;
        for (int w = tid; w < W; w += THREADS) d += __popc(sm[w]);
        d = block_sum<THREADS>(d, red);
        f = block_sum<THREADS>(f, red);
        if (tid == 0) {
            D[row - rb] = d;
            F[row - rb] = f;
            atomicMax((long long*)&scalars[SC_MAXD], (long long)d);
        }
    }
}

// =========================================================================================
// step 1, kernel 3 (fill): block per tile row.  Rebuilds the window bitmap, ranks every kept
// product by its tile column (prefix popcount = position in the ascending C' column list),
// counts pairs per C' tile, scans, emits C' (row, col, pair offset) and finally places the pairs
// in ascending-k order: the A tiles of the row are visited in order and, within one A tile, all
// B tiles have distinct columns, so `off[rank]++` needs no atomics.
// Shared memory: bitmap[Wmax] | prefix[Wmax] | cnt[cap]   (cnt falls back to global scratch when
// a row has more than `cap` C' tiles).
// =========================================================================================
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_step1_fill(int rb, int re, const int32_t* __restrict__ Arp, const int32_t* __restrict__ Acol,
             const uint16_t* __restrict__ AcolOcc, const int32_t* __restrict__ Brp,
             const int32_t* __restrict__ Bcol, const uint16_t* __restrict__ BrowOcc,
             const int2* __restrict__ win, int keep_empty, int Wmax, int cap,
             unsigned* __restrict__ gcnt, size_t gcnt_stride,
             const int64_t* __restrict__ c_row_ptr, const int64_t* __restrict__ pair_row_ptr,
             int32_t* __restrict__ c_tile_row, int32_t* __restrict__ c_tile_col,
             int64_t* __restrict__ pair_ptr, int32_t* __restrict__ pairs_a, int32_t* __restrict__ pairs_b)
{
    extern __shared__ unsigned sm[];
    __shared__ unsigned red[THREADS / 32 + 1];
    constexpr int NW = THREADS / 32;
    unsigned* bitmap = sm;
    unsigned* prefix = sm + Wmax;
    unsigned* cnt_sm = sm + 2 * (size_t)Wmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int row = rb + blockIdx.x; row < re; row += gridDim.x) {
        const int64_t cbase = c_row_ptr[row - rb];
        const int D = (int)(c_row_ptr[row - rb + 1] - cbase);
        if (D == 0) continue;
        const int64_t pbase = pair_row_ptr[row - rb];
        const int2 wnd = win[row - rb];
        const int base = wnd.x & ~31;
        const int W = ((wnd.y - base) >> 5) + 1;
        unsigned* cnt = (D <= cap) ? cnt_sm : gcnt + (size_t)blockIdx.x * gcnt_stride;
        for (int w = tid; w < W; w += THREADS) bitmap[w] = 0;
        for (int i = tid; i < D; i += THREADS) cnt[i] = 0;
        __syncthreads();
        const int as = Arp[row], ae = Arp[row + 1];
        // pass A: structure
        for (int p = as + warp; p < ae; p += NW) {
            const int k = Acol[p];
            const unsigned aocc = keep_empty ? 0xFFFFu : AcolOcc[p];
            const int bs = Brp[k], be = Brp[k + 1];
            for (int q = bs + lane; q < be; q += 32)
                if (aocc & BrowOcc[q]) {
                    int j = Bcol[q] - base;
                    atomicOr(&bitmap[j >> 5], 1u << (j & 31));
                }
        }
        __syncthreads();
        block_scan_exclusive<THREADS>(prefix, W, [&](int i) { return (unsigned)__popc(bitmap[i]); }, red);
        // pass B: pairs per C' tile
        for (int p = as + warp; p < ae; p += NW) {
            const int k = Acol[p];
            const unsigned aocc = keep_empty ? 0xFFFFu : AcolOcc[p];
            const int bs = Brp[k], be = Brp[k + 1];
            for (int q = bs + lane; q < be; q += 32)
                if (aocc & BrowOcc[q]) {
                    int j = Bcol[q] - base;
                    int w = j >> 5;
                    unsigned rank = prefix[w] + __popc(bitmap[w] & ((1u << (j & 31)) - 1u));
                    atomicAdd(&cnt[rank], 1u);
                }
        }
        __syncthreads();
        block_scan_exclusive<THREADS>(cnt, D, [&](int i) { return cnt[i]; }, red);
        // emit the C' tiles of this row (columns ascending) with their pair offsets
        for (int w = tid; w < W; w += THREADS) {
            unsigned m = bitmap[w];
            unsigned r = prefix[w];
            while (m) {
                int b = __ffs(m) - 1;
                m &= m - 1;
                c_tile_row[cbase + r] = row;
                c_tile_col[cbase + r] = base + w * 32 + b;
                pair_ptr[cbase + r] = pbase + cnt[r];
                ++r;
            }
        }
        __syncthreads();
        // pass C: ordered placement, A tiles in ascending k
        for (int p = as; p < ae; ++p) {
            const int k = Acol[p];
            const unsigned aocc = keep_empty ? 0xFFFFu : AcolOcc[p];
            const int bs = Brp[k], be = Brp[k + 1];
            for (int q = bs + tid; q < be; q += THREADS)
                if (aocc & BrowOcc[q]) {
                    int j = Bcol[q] - base;
                    int w = j >> 5;
                    unsigned rank = prefix[w] + __popc(bitmap[w] & ((1u << (j & 31)) - 1u));
                    unsigned pos = cnt[rank]++;
                    pairs_a[pbase + pos] = p;
                    pairs_b[pbase + pos] = q;
                }
            __syncthreads();
        }
    }
}

// =========================================================================================
// step 2, kernel 1: one thread per C' tile: mask of the tile = OR over its pairs of the boolean
// product of the two 16x16 bit matrices, computed as  Cmask[r] |= OR_{k in Amask[r]} Bmask[k].
// =========================================================================================
__global__ void __launch_bounds__(128)
k_step2_masks(int64_t n_tiles, const int64_t* __restrict__ pair_ptr, const int32_t* __restrict__ pairs_a,
              const int32_t* __restrict__ pairs_b, const uint16_t* __restrict__ Amasks,
              const uint16_t* __restrict__ Bmasks, uint16_t* __restrict__ Cmasks,
              int64_t* __restrict__ c_tile_nnz)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    unsigned acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t ps = pair_ptr[t], pe = pair_ptr[t + 1];
    for (int64_t pp = ps; pp < pe; ++pp) {
        const int a = pairs_a[pp], b = pairs_b[pp];
        const uint4* am4 = reinterpret_cast<const uint4*>(Amasks + (size_t)a * 16);
        const uint4 x = am4[0], y = am4[1];
        const unsigned aw[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
        const uint16_t* __restrict__ bm = Bmasks + (size_t)b * 16;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            unsigned m = (aw[r >> 1] >> ((r & 1) * 16)) & 0xFFFFu;
            unsigned o = 0;
            while (m) {
                int k = __ffs(m) - 1;
                m &= m - 1;
                o |= bm[k];
            }
            acc[r >> 1] |= o << ((r & 1) * 16);
        }
    }
    uint4* out = reinterpret_cast<uint4*>(Cmasks + (size_t)t * 16);
    out[0] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    out[1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
    int nnz = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) nnz += __popc(acc[i]);
    c_tile_nnz[t] = nnz;
}

// =========================================================================================
// step 2, kernel 2: one thread per C' tile: Ctiles_rowColIdx ((r<<4)|c per nonzero, row-major)
// and the first tile of every step-3 block (a tile holds <= 256 = PEM_S3_ENTRIES nonzeros, so it
// covers at most one block boundary).
// =========================================================================================
__global__ void __launch_bounds__(128)
k_step2_rowcolidx(int64_t n_tiles, const uint16_t* __restrict__ Cmasks, const int64_t* __restrict__ c_tile_nnz_ptr,
                  uint8_t* __restrict__ row_col_idx, int32_t* __restrict__ blk_tile)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint4* m4 = reinterpret_cast<const uint4*>(Cmasks + (size_t)t * 16);
    const uint4 x = m4[0], y = m4[1];
    const unsigned w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    int64_t off = c_tile_nnz_ptr[t];
    const int64_t end = c_tile_nnz_ptr[t + 1];
    if (end > off) {
        int64_t bnd = (off + PEM_S3_ENTRIES - 1) / PEM_S3_ENTRIES * PEM_S3_ENTRIES;
        if (bnd < end) blk_tile[bnd / PEM_S3_ENTRIES] = (int32_t)t;
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        unsigned m = (w[r >> 1] >> ((r & 1) * 16)) & 0xFFFFu;
        while (m) {
            int c = __ffs(m) - 1;
            m &= m - 1;
            row_col_idx[off++] = (uint8_t)((r << 4) | c);
        }
    }
}

// =========================================================================================
// step 3: one thread per C nonzero.  The block covers PEM_S3_ENTRIES consecutive nonzeros; a
// thread finds its tile by binary search in the tile offset array between the block's first tile
// and the next block's first tile, then walks the tile's pair list in order:
//   m = Amask[r] & BmaskT[c];  for each k in m (ascending):  acc = fma(a_rk, b_kc, acc)
// Same product order as the reference (:648-656) and as the host oracle; C is written once.
// =========================================================================================
__global__ void __launch_bounds__(PEM_S3_ENTRIES)
k_step3_numeric(int64_t nnz, int64_t n_tiles, const int32_t* __restrict__ blk_tile,
                const int64_t* __restrict__ c_tile_nnz_ptr, const uint8_t* __restrict__ row_col_idx,
                const int64_t* __restrict__ pair_ptr, const int32_t* __restrict__ pairs_a,
                const int32_t* __restrict__ pairs_b,
                const uint32_t* __restrict__ A_off, const double* __restrict__ A_vals,
                const uint16_t* __restrict__ A_masks, const uint8_t* __restrict__ A_rowptr,
                const uint32_t* __restrict__ B_off, const double* __restrict__ B_vals,
                const uint16_t* __restrict__ B_masks, const uint8_t* __restrict__ B_rowptr,
                const uint16_t* __restrict__ B_masks_t, double* __restrict__ C_vals)
{
    const int64_t n = (int64_t)blockIdx.x * PEM_S3_ENTRIES + threadIdx.x;
    if (n >= nnz) return;
    // last tile t in [lo, hi] with c_tile_nnz_ptr[t] <= n
    int64_t lo = blk_tile[blockIdx.x];
    int64_t hi = ((int64_t)(blockIdx.x + 1) * PEM_S3_ENTRIES < nnz) ? blk_tile[blockIdx.x + 1] : n_tiles - 1;
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if (c_tile_nnz_ptr[mid] <= n) lo = mid; else hi = mid - 1;
    }
    const int64_t t = lo;
    const unsigned rc = row_col_idx[n];
    const unsigned r = rc >> 4, c = rc & 15u;
    const unsigned below_c = (1u << c) - 1u;
    double acc = 0.0;
    const int64_t ps = pair_ptr[t], pe = pair_ptr[t + 1];
    for (int64_t pp = ps; pp < pe; ++pp) {
        const int a = pairs_a[pp], b = pairs_b[pp];
        const unsigned am = A_masks[(size_t)a * 16 + r];
        unsigned m = am & B_masks_t[(size_t)b * 16 + c];
        if (m) {
            const double* __restrict__ av = A_vals + A_off[a] + A_rowptr[(size_t)a * 16 + r];
            const double* __restrict__ bv = B_vals + B_off[b];
            do {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const int ao = __popc(am & ((1u << k) - 1u));
                const int bo = B_rowptr[(size_t)b * 16 + k] + __popc(B_masks[(size_t)b * 16 + k] & below_c);
                acc = fma(av[ao], bv[bo], acc);
            } while (m);
        }
    }
    C_vals[n] = acc;
}

__global__ void k_set_last_i64(int64_t* p, int64_t idx, int64_t v) { p[idx] = v; }

}  // namespace

// =========================================================================================
// host side
// =========================================================================================
static int check_operands(pem_ctx* ctx, const pem_tiled* A, const pem_tiled* B, int rb, int re)
{
    if (!ctx || !A || !B) return PEM_ERR_ARG;
    if (A->cols != B->rows) return ctx->fail(PEM_ERR_ARG, "inner dimensions differ (A.cols != B.rows)");
    if (rb < 0 || re < rb || re > A->tile_rows) return ctx->fail(PEM_ERR_ARG, "tile-row panel out of range");
    return PEM_OK;
}

extern "C" {

int pem_step1_symbolic(pem_ctx* ctx, const pem_tiled* A, const pem_tiled* B,
                       int32_t rb, int32_t re, pem_result** out)
{
    if (!out) return PEM_ERR_ARG;
    *out = nullptr;
    PEM_TRY(check_operands(ctx, A, B, rb, re));
    PEM_CK(cudaSetDevice(ctx->device));
    pem_result* C = new pem_result();
    C->rb = rb; C->re = re; C->rows = A->rows; C->cols = B->cols; C->tile_cols = B->tile_cols;
    auto fail = [&](int rc) { pem_result_free(ctx, C); return rc; };
#define S_TRY(expr) do { int rc_ = (expr); if (rc_ != PEM_OK) return fail(rc_); } while (0)
#define S_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx->fail_cuda(e_, #call, __FILE__, __LINE__)); } while (0)
    const int nrows = re - rb;
    S_TRY(pem_alloc(ctx, &C->row_ptr, (size_t)nrows + 1));
    int64_t* pair_row_ptr = nullptr;
    int2* win = nullptr;
    S_TRY(pem_alloc(ctx, &pair_row_ptr, (size_t)nrows + 1));
    S_TRY(pem_alloc(ctx, &win, (size_t)nrows));
    S_CK(cudaMemsetAsync(ctx->d_scalars, 0, PEM_NSCALARS * sizeof(int64_t), ctx->stream));
    S_CK(cudaMemsetAsync(C->row_ptr, 0, ((size_t)nrows + 1) * 8, ctx->stream));
    S_CK(cudaMemsetAsync(pair_row_ptr, 0, ((size_t)nrows + 1) * 8, ctx->stream));
    auto cleanup_tmp = [&]() { pem_free(ctx, pair_row_ptr); pem_free(ctx, win); };

    int64_t maxwin = 0, maxd = 0;
    if (nrows > 0 && A->tiles > 0 && B->tiles > 0) {
        k_row_window<<<pem_div_up((int64_t)nrows * 32, 256), 256, 0, ctx->stream>>>(
            rb, re, A->tile_row_ptr, A->tile_col_idx, B->tile_row_ptr, B->tile_col_idx, win, ctx->d_scalars);
        ++ctx->launches;
        S_CK(cudaGetLastError());
        S_CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, PEM_NSCALARS * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        S_CK(cudaStreamSynchronize(ctx->stream));
        maxwin = ctx->h_scalars[SC_MAXWIN];
        C->tile_products = ctx->h_scalars[SC_SUMP];
    }
    constexpr int TH = 128;
    if (maxwin > 0) {
        size_t smem_count = (size_t)maxwin * 4;
        if (smem_count > (size_t)ctx->smem_optin) {
            cleanup_tmp();
            return fail(ctx->fail(PEM_ERR_LIMIT, "step 1: column window of a tile row exceeds the shared-memory bitmap"));
        }
        S_CK(cudaFuncSetAttribute(k_step1_count<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
        int grid = (int)std::min<int64_t>(nrows, (int64_t)ctx->sm_count * 64);
        k_step1_count<TH><<<grid, TH, smem_count, ctx->stream>>>(
            rb, re, A->tile_row_ptr, A->tile_col_idx, A->col_occ, B->tile_row_ptr, B->tile_col_idx, B->row_occ,
            win, ctx->opt_keep_empty, C->row_ptr, pair_row_ptr, ctx->d_scalars);
        ++ctx->launches;
        S_CK(cudaGetLastError());
        S_TRY(pem_scan_exclusive_i64(ctx, C->row_ptr, (int64_t)nrows + 1));
        S_TRY(pem_scan_exclusive_i64(ctx, pair_row_ptr, (int64_t)nrows + 1));
        S_CK(cudaMemcpyAsync(&ctx->d_scalars[SC_T0], C->row_ptr + nrows, 8, cudaMemcpyDeviceToDevice, ctx->stream));
        S_CK(cudaMemcpyAsync(&ctx->d_scalars[SC_T1], pair_row_ptr + nrows, 8, cudaMemcpyDeviceToDevice, ctx->stream));
        S_CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, PEM_NSCALARS * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        S_CK(cudaStreamSynchronize(ctx->stream));
        C->tiles = ctx->h_scalars[SC_T0];
        C->pairs = ctx->h_scalars[SC_T1];
        maxd = ctx->h_scalars[SC_MAXD];
    }
    S_TRY(pem_alloc(ctx, &C->tile_row, (size_t)C->tiles));
    S_TRY(pem_alloc(ctx, &C->tile_col, (size_t)C->tiles));
    S_TRY(pem_alloc(ctx, &C->pair_ptr, (size_t)C->tiles + 1));
    S_TRY(pem_alloc(ctx, &C->pairs_a, (size_t)C->pairs));
    S_TRY(pem_alloc(ctx, &C->pairs_b, (size_t)C->pairs));
    if (C->tiles > 0) {
        const int cap_default = 4096;
        size_t smem_fill = (size_t)maxwin * 8 + (size_t)cap_default * 4;
        int cap = cap_default;
        if (smem_fill > (size_t)ctx->smem_optin) {
            cap = 0;
            smem_fill = (size_t)maxwin * 8;
            if (smem_fill > (size_t)ctx->smem_optin) {
                cleanup_tmp();
                return fail(ctx->fail(PEM_ERR_LIMIT, "step 1: column window of a tile row exceeds the shared-memory bitmap"));
            }
        }
        int grid = (int)std::min<int64_t>(nrows, (int64_t)ctx->sm_count * 8);
        unsigned* gcnt = nullptr;
        size_t stride = 0;
        if (maxd > cap) {
            stride = (size_t)maxd;
            S_TRY(pem_alloc(ctx, &gcnt, stride * (size_t)grid));
        }
        S_CK(cudaFuncSetAttribute(k_step1_fill<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fill));
        k_step1_fill<TH><<<grid, TH, smem_fill, ctx->stream>>>(
            rb, re, A->tile_row_ptr, A->tile_col_idx, A->col_occ, B->tile_row_ptr, B->tile_col_idx, B->row_occ,
            win, ctx->opt_keep_empty, (int)maxwin, cap, gcnt, stride, C->row_ptr, pair_row_ptr,
            C->tile_row, C->tile_col, C->pair_ptr, C->pairs_a, C->pairs_b);
        ++ctx->launches;
        S_CK(cudaGetLastError());
        pem_free(ctx, gcnt);
    }
    k_set_last_i64<<<1, 1, 0, ctx->stream>>>(C->pair_ptr, C->tiles, C->pairs);
    ++ctx->launches;
    S_CK(cudaGetLastError());
    cleanup_tmp();
    C->stage = 1;
    *out = C;
    return PEM_OK;
}

int pem_step2_symbolic(pem_ctx* ctx, const pem_tiled* A, const pem_tiled* B, pem_result* C)
{
    if (!ctx || !A || !B || !C) return PEM_ERR_ARG;
    if (C->stage != 1) return ctx->fail(PEM_ERR_ARG, "step 2 needs a result fresh from step 1");
    PEM_CK(cudaSetDevice(ctx->device));
    PEM_TRY(pem_alloc(ctx, &C->masks, (size_t)C->tiles * 16));
    PEM_TRY(pem_alloc(ctx, &C->tile_nnz_ptr, (size_t)C->tiles + 1));
    if (C->tiles > 0) {
        k_step2_masks<<<pem_div_up(C->tiles, 128), 128, 0, ctx->stream>>>(
            C->tiles, C->pair_ptr, C->pairs_a, C->pairs_b, A->masks, B->masks, C->masks, C->tile_nnz_ptr);
        PEM_LAUNCHED();
    }
    k_set_last_i64<<<1, 1, 0, ctx->stream>>>(C->tile_nnz_ptr, C->tiles, 0);
    PEM_LAUNCHED();
    PEM_TRY(pem_scan_exclusive_i64(ctx, C->tile_nnz_ptr, C->tiles + 1));
    PEM_CK(cudaMemcpyAsync(ctx->h_scalars, C->tile_nnz_ptr + C->tiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PEM_CK(cudaStreamSynchronize(ctx->stream));
    C->nnz = ctx->h_scalars[0];
    PEM_TRY(pem_alloc(ctx, &C->row_col_idx, (size_t)C->nnz));
    int64_t nblk = (C->nnz + PEM_S3_ENTRIES - 1) / PEM_S3_ENTRIES;
    PEM_TRY(pem_alloc(ctx, &C->blk_tile, (size_t)nblk + 1));
    if (C->tiles > 0) {
        k_step2_rowcolidx<<<pem_div_up(C->tiles, 128), 128, 0, ctx->stream>>>(
            C->tiles, C->masks, C->tile_nnz_ptr, C->row_col_idx, C->blk_tile);
        PEM_LAUNCHED();
    }
    C->stage = 2;
    return PEM_OK;
}

int pem_step3_numeric(pem_ctx* ctx, const pem_tiled* A, const pem_tiled* B, pem_result* C)
{
    if (!ctx || !A || !B || !C) return PEM_ERR_ARG;
    if (C->stage != 2) return ctx->fail(PEM_ERR_ARG, "step 3 needs a result fresh from step 2");
    PEM_CK(cudaSetDevice(ctx->device));
    PEM_TRY(pem_alloc(ctx, &C->vals, (size_t)C->nnz));
    if (C->nnz > 0) {
        int64_t nblk = (C->nnz + PEM_S3_ENTRIES - 1) / PEM_S3_ENTRIES;
        if (nblk > 0x7fffffffLL) return ctx->fail(PEM_ERR_LIMIT, "C has more than 2^39 nonzeros");
        k_step3_numeric<<<(unsigned)nblk, PEM_S3_ENTRIES, 0, ctx->stream>>>(
            C->nnz, C->tiles, C->blk_tile, C->tile_nnz_ptr, C->row_col_idx, C->pair_ptr, C->pairs_a, C->pairs_b,
            A->tile_nnz_ptr, A->vals, A->masks, A->row_ptr, B->tile_nnz_ptr, B->vals, B->masks, B->row_ptr,
            B->masks_t, C->vals);
        PEM_LAUNCHED();
    }
    C->stage = 3;
    return PEM_OK;
}

int pem_spgemm_panel(pem_ctx* ctx, const pem_tiled* A, const pem_tiled* B,
                     int32_t rb, int32_t re, pem_result** out, pem_times* times)
{
    if (!out) return PEM_ERR_ARG;
    *out = nullptr;
    PEM_TRY(check_operands(ctx, A, B, rb, re));
    PEM_CK(cudaSetDevice(ctx->device));
    PEM_CK(cudaStreamSynchronize(ctx->stream));
    auto w0 = std::chrono::high_resolution_clock::now();
    pem_result* C = nullptr;
    PEM_CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    PEM_TRY(pem_step1_symbolic(ctx, A, B, rb, re, &C));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    int rc = pem_step2_symbolic(ctx, A, B, C);
    cudaEventRecord(ctx->ev[4], ctx->stream);
    if (rc == PEM_OK) rc = pem_step3_numeric(ctx, A, B, C);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    if (rc == PEM_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        rc = ctx->fail_cuda(cudaGetLastError(), "cudaStreamSynchronize after step 3", __FILE__, __LINE__);
    if (rc != PEM_OK) { pem_result_free(ctx, C); return rc; }
    auto w1 = std::chrono::high_resolution_clock::now();
    if (times) {
        float s1 = 0, s2 = 0, s3 = 0;
        cudaEventElapsedTime(&s1, ctx->ev[2], ctx->ev[3]);
        cudaEventElapsedTime(&s2, ctx->ev[3], ctx->ev[4]);
        cudaEventElapsedTime(&s3, ctx->ev[4], ctx->ev[5]);
        times->step1_ms = s1; times->step2_ms = s2; times->step3_ms = s3;
        times->kernel_ms = (double)s1 + s2 + s3;
        times->total_ms = std::chrono::duration<double, std::milli>(w1 - w0).count();
        times->malloc_ms = times->total_ms - times->kernel_ms;
    }
    *out = C;
    return PEM_OK;
}

int pem_spgemm(pem_ctx* ctx, const pem_tiled* A, const pem_tiled* B, pem_result** C, pem_times* times)
{
    if (!A) return PEM_ERR_ARG;
    return pem_spgemm_panel(ctx, A, B, 0, A->tile_rows, C, times);
}

}  // extern "C"
