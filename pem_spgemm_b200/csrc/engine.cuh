// Internal definitions shared by the engine's translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "pemspgemm.h"

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
enum { PEM_NSCALARS = 16, PEM_NEVENTS = 8 };

struct pem_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    std::string err;
    int64_t launches = 0;        // kernels of this library launched on the stream
    int opt_keep_empty = 0;      // PEM_OPT_KEEP_EMPTY_TILES
    int opt_step1_path = 0;      // PEM_OPT_STEP1_PATH
    int sm_count = 148;
    int smem_optin = 227 * 1024; // max dynamic shared memory per block
    int64_t* h_scalars = nullptr; // pinned, PEM_NSCALARS entries: size read-backs
    int64_t* d_scalars = nullptr; // device mirror the kernels reduce into
    cudaEvent_t ev[PEM_NEVENTS] = {};

    int fail(int code, const std::string& msg) { err = msg; return code; }
    int fail_cuda(cudaError_t e, const char* what, const char* file, int line)
    {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
        err = buf;
        (void)cudaGetLastError();
        return PEM_ERR_CUDA;
    }
};

#define PEM_CK(call)                                                                     \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) return ctx->fail_cuda(e__, #call, __FILE__, __LINE__);   \
    } while (0)

// kernel launch check + launch accounting
#define PEM_LAUNCHED()                                                                   \
    do {                                                                                 \
        ++ctx->launches;                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) return ctx->fail_cuda(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define PEM_TRY(expr)                    \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != PEM_OK) return rc__; \
    } while (0)

// stream-ordered allocation from the context's pool (never returns a null pointer for n == 0)
template <class T>
static inline int pem_alloc(pem_ctx* ctx, T** p, size_t n)
{
    size_t bytes = (n ? n : 1) * sizeof(T);
    PEM_CK(cudaMallocFromPoolAsync((void**)p, bytes, ctx->pool, ctx->stream));
    return PEM_OK;
}
template <class T>
static inline void pem_free(pem_ctx* ctx, T*& p)
{
    if (p) cudaFreeAsync((void*)p, ctx->stream);
    p = nullptr;
}

static inline int pem_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------
// tiled matrix (arrays named after the reference's)
// ---------------------------------------------------------------------------------------
struct pem_tiled {
    int32_t rows = 0, cols = 0;
    int64_t nnz = 0;
    int32_t tile_rows = 0, tile_cols = 0, tiles = 0;
    double* vals = nullptr;           // [nnz] tile-major, row-major inside a tile
    uint32_t* tile_nnz_ptr = nullptr; // [tiles+1]
    uint16_t* masks = nullptr;        // [tiles*16] row masks
    uint16_t* masks_t = nullptr;      // [tiles*16] column masks
    uint8_t* row_ptr = nullptr;       // [tiles*16]
    int32_t* tile_row_ptr = nullptr;  // [tile_rows+1]
    int32_t* tile_col_idx = nullptr;  // [tiles]
    int32_t* tile_row_idx = nullptr;  // [tiles]
    uint16_t* col_occ = nullptr;      // [tiles]
    uint16_t* row_occ = nullptr;      // [tiles]
};

// ---------------------------------------------------------------------------------------
// result: C or a tile-row panel of C
// ---------------------------------------------------------------------------------------
struct pem_result {
    int32_t rb = 0, re = 0;           // panel of A' tile rows
    int32_t rows = 0, cols = 0;       // shape of the full C
    int32_t tile_cols = 0;
    int64_t tiles = 0, pairs = 0, nnz = 0, tile_products = 0;
    int stage = 0;                    // 1, 2, 3 = last completed step
    int64_t* row_ptr = nullptr;       // [re-rb+1]
    int32_t* tile_row = nullptr;      // [tiles]
    int32_t* tile_col = nullptr;      // [tiles]
    int64_t* pair_ptr = nullptr;      // [tiles+1]
    int32_t* pairs_a = nullptr;       // [pairs]
    int32_t* pairs_b = nullptr;       // [pairs]
    uint16_t* masks = nullptr;        // [tiles*16]
    int64_t* tile_nnz_ptr = nullptr;  // [tiles+1]
    uint8_t* row_col_idx = nullptr;   // [nnz]
    double* vals = nullptr;           // [nnz]
    int32_t* blk_tile = nullptr;      // [ceil(nnz/PEM_S3_ENTRIES)+1] first tile of each step-3 block
};

enum { PEM_S3_ENTRIES = 256 };  // C entries per step-3 thread block (>= 256 so a tile spans <= 2 blocks)

// scalars slots in ctx->d_scalars / h_scalars
enum {
    SC_ERR = 0,       // error flags from kernels
    SC_COUNT = 1,     // generic count (tiles found by the census)
    SC_MAXWIN = 2,    // step 1: max window words over the panel's rows
    SC_MAXP = 3,      // step 1: max tile products of a row
    SC_MAXD = 4,      // step 1: max C' tiles of a row
    SC_SUMP = 5,      // step 1: total tile products
    SC_T0 = 6, SC_T1 = 7, SC_T2 = 8
};

// step entry points implemented in spgemm.cu / convert.cu / export.cu
int pem_scan_exclusive_i64(pem_ctx* ctx, int64_t* d_inout, int64_t n);  // in place, n elements
