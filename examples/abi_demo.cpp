// Minimal C++ client of the C ABI (include/pemspgemm.h): what the reference's main would call
// in place of its inline conversion / SpGEMM loop / export.
//   abi_demo <grid>     C = A^2 for the 2-D 5-point Laplacian on a grid x grid mesh,
//                       then C = A * A^T through pem_tiled_transpose; prints nnz and checksums.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pemspgemm.h"

#define CK(call)                                                                           \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != PEM_OK) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ctx ? pem_last_error(ctx) : ""); \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

int main(int argc, char** argv)
{
    const int g = argc > 1 ? atoi(argv[1]) : 64;
    const int n = g * g;
    std::vector<int32_t> I, J;
    std::vector<double> V;
    for (int y = 0; y < g; ++y)
        for (int x = 0; x < g; ++x) {
            const int i = y * g + x;
            auto put = [&](int j, double v) { I.push_back(i); J.push_back(j); V.push_back(v); };
            put(i, 4.0);
            if (x > 0) put(i - 1, -1.0);
            if (x < g - 1) put(i + 1, -1.0);
            if (y > 0) put(i - g, -1.0);
            if (y < g - 1) put(i + g, -1.0);
        }
    pem_ctx* ctx = nullptr;
    CK(pem_ctx_create(&ctx, 0));
    pem_tiled *A = nullptr, *At = nullptr;
    pem_times tc = {};
    CK(pem_convert_coo(ctx, n, n, (int64_t)I.size(), I.data(), J.data(), V.data(), 0, &A, &tc));
    uint64_t flop = 0;
    CK(pem_count_flop(ctx, A, A, &flop));
    pem_result* C = nullptr;
    pem_times t = {};
    CK(pem_spgemm(ctx, A, A, &C, &t));
    pem_result_info ci;
    CK(pem_result_info_get(C, &ci));
    double sum = 0, asum = 0;
    CK(pem_result_checksum(ctx, C, &sum, &asum));
    std::vector<int32_t> r((size_t)ci.nnz), c((size_t)ci.nnz);
    std::vector<double> v((size_t)ci.nnz);
    CK(pem_result_to_coo(ctx, C, r.data(), c.data(), v.data()));
    printf("A^2: n=%d nnzA=%zu flop=%llu C_tiles=%lld C_nnz=%lld sum=%.1f abs_sum=%.1f C[0,0]=%.1f step1/2/3=%.3f/%.3f/%.3f ms\n",
           n, I.size(), (unsigned long long)flop, (long long)ci.tiles, (long long)ci.nnz, sum, asum,
           ci.nnz ? v[0] : 0.0, t.step1_ms, t.step2_ms, t.step3_ms);
    pem_result_free(ctx, C);
    CK(pem_tiled_transpose(ctx, A, &At));
    CK(pem_spgemm(ctx, A, At, &C, nullptr));
    CK(pem_result_info_get(C, &ci));
    CK(pem_result_checksum(ctx, C, &sum, &asum));
    printf("A*A^T: C_nnz=%lld sum=%.1f abs_sum=%.1f\n", (long long)ci.nnz, sum, asum);
    pem_result_free(ctx, C);
    pem_tiled_free(ctx, At);
    pem_tiled_free(ctx, A);
    pem_ctx_destroy(ctx);
    return 0;
}
