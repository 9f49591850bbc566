// Micro-benchmark: how many small 1-D bulk copies (cp.async.bulk global -> shared, mbarrier complete_tx) per
// microsecond an SM sustains, against LDGSTS (cp.async 16 B) moving the same bytes; STAGES rounds in flight,
// source offsets computed (no dependent load).  Decides the staging engine of step 3's staged kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_rate bulk_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

constexpr int STAGES = 4;

template <int MODE>   // 0 = bulk (one lane per blob), 1 = LDGSTS (16 lanes per blob)
__global__ void __launch_bounds__(256) k_rate(const char* __restrict__ src, uint32_t nunits, int blob_bytes,
                                               int rounds, int blobs_per_round, int issue_lanes, unsigned long long* sink)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) unsigned long long bar[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(issue_lanes));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const size_t region = (size_t)blobs_per_round * blob_bytes;
    unsigned long long acc = 0;
    auto issue = [&](int it) {
        const int s = it % STAGES;
        char* dst = smem + s * region;
        const uint32_t seed = (blockIdx.x * 1000003u + it) * 4099u;
        if (MODE == 0) {
            if (tid < issue_lanes) {
                int mine = 0;
                for (int b = tid; b < blobs_per_round; b += issue_lanes) ++mine;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(mine * blob_bytes) : "memory");
                for (int b = tid; b < blobs_per_round; b += issue_lanes) {
                    const uint32_t o = mix(seed + b) % nunits;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(dst + (size_t)b * blob_bytes)), "l"(src + (size_t)o * 16), "r"(blob_bytes), "r"(smem_u32(&bar[s])) : "memory");
                }
            }
        } else {
            const int units = blob_bytes / 16;
            for (int b = tid >> 4; b < blobs_per_round; b += blockDim.x >> 4) {
                const uint32_t o = mix(seed + b) % nunits;
                const char* g = src + (size_t)o * 16;
                for (int u = tid & 15; u < units; u += 16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + (size_t)b * blob_bytes + u * 16)), "l"(g + u * 16) : "memory");
            }
            asm volatile("cp.async.commit_group;");
        }
    };
    for (int it = 0; it < STAGES - 1 && it < rounds; ++it) issue(it);
    for (int it = 0; it < rounds; ++it) {
        if (it + STAGES - 1 < rounds) issue(it + STAGES - 1);
        else if (MODE == 1) asm volatile("cp.async.commit_group;");
        const int s = it % STAGES;
        if (MODE == 0) {
            uint32_t ok = 0;
            const uint32_t parity = (it / STAGES) & 1;
            while (!ok) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
            }
        } else {
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
            __syncthreads();
        }
        acc += *reinterpret_cast<unsigned long long*>(smem + s * region + ((tid * 16) % region));
        __syncthreads();
    }
    if (acc == 0x1234567) sink[0] = acc;
}

int main(int argc, char** argv)
{
    unsigned long long* sink; cudaMalloc(&sink, 8);
    const int rounds = 400;
    for (size_t mb : {64, 4096}) {
        const size_t src_bytes = mb << 20;
        char* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
        const uint32_t nunits = (uint32_t)((src_bytes - 4096) / 16);
        for (int bps : {1, 2}) {
            const int blocks = 148 * bps;
            for (int blob_bytes : {64, 208, 416}) {
                for (int blobs_per_round : {64, 128}) {
                    const size_t smem = (size_t)blobs_per_round * blob_bytes * STAGES;
                    if (smem * bps > 200 * 1024) continue;
                    cudaFuncSetAttribute(k_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    for (int mode = 0; mode < 2; ++mode) {
                        for (int lanes : {32, 64}) {
                            if (mode == 1 && lanes != 32) continue;
                            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                            for (int rep = 0; rep < 2; ++rep) {
                                cudaEventRecord(e0);
                                if (mode == 0) k_rate<0><<<blocks, 256, smem>>>(src, nunits, blob_bytes, rounds, blobs_per_round, lanes, sink);
                                else k_rate<1><<<blocks, 256, smem>>>(src, nunits, blob_bytes, rounds, blobs_per_round, lanes, sink);
                                cudaEventRecord(e1); cudaEventSynchronize(e1);
                            }
                            float ms; cudaEventElapsedTime(&ms, e0, e1);
                            const double copies = (double)blocks * rounds * blobs_per_round;
                            printf("src %4zu MB, %d blk/SM, %s blob %4d B x %3d per round, %3d issue lanes: %.3f ms, %.1f copies/us/SM, %.1f GB/s, err=%d\n",
                                   mb, bps, mode == 0 ? "bulk  " : "ldgsts", blob_bytes, blobs_per_round, lanes, ms, copies / (ms * 1e3) / 148.0,
                                   copies * blob_bytes / (ms * 1e6), (int)cudaGetLastError());
                        }
                    }
                }
            }
        }
        cudaFree(src);
    }
    return 0;
}
